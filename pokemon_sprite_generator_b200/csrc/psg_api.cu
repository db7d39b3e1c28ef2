// Library-level entry points: version, last-error string, device probe.
#include "psg_common.cuh"
#include <stdarg.h>

static thread_local char g_last_error[512] = "";
long long g_psg_launch_count = 0;

void psg_set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_last_error, sizeof(g_last_error), fmt, ap);
  va_end(ap);
}

extern "C" {

int psg_version() { return 100; }  // 0.1.0

const char* psg_last_error() { return g_last_error; }

// Number of kernels this library has launched since load (or since the last reset); bench.py reports it.
long long psg_launch_count(int reset) {
  long long v = g_psg_launch_count;
  if (reset) g_psg_launch_count = 0;
  return v;
}

// Returns 0 when the current device is compute capability 10.x (the only target of this library).
int psg_check_device() {
  int dev = 0;
  cudaError_t e = cudaGetDevice(&dev);
  if (e != cudaSuccess) { psg_set_error("psg_check_device: %s", cudaGetErrorString(e)); return PSG_ERR_CUDA; }
  int major = 0, minor = 0;
  cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, dev);
  cudaDeviceGetAttribute(&minor, cudaDevAttrComputeCapabilityMinor, dev);
  if (major != 10) {
    psg_set_error("psg_check_device: device is sm_%d%d; this library is built for sm_100a only", major, minor);
    return PSG_ERR_UNSUPPORTED;
  }
  return PSG_OK;
}

}  // extern "C"
