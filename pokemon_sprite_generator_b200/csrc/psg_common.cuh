// Shared helpers for the psg_b200 C-ABI library (sm_100a only).
//
// Conventions: every entry point is `extern "C" int psg_*(..., void* stream)`,
// returns 0 or a negative error code, never allocates, never synchronises; the caller owns
// all memory.  `psg_last_error()` returns a thread-local description of the last failure.
#pragma once

#include <cuda_runtime.h>
#include <cuda_bf16.h>
#include <stdint.h>
#include <stdio.h>

#define PSG_OK 0
#define PSG_ERR_INVALID -1
#define PSG_ERR_CUDA -2
#define PSG_ERR_UNSUPPORTED -3

void psg_set_error(const char* fmt, ...);
extern long long g_psg_launch_count;  // kernels launched by this library (host-side counter, see psg_launch_count)

#define PSG_CHECK_ARG(cond, ...)                     \
  do {                                               \
    if (!(cond)) {                                   \
      psg_set_error(__VA_ARGS__);                    \
      return PSG_ERR_INVALID;                        \
    }                                                \
  } while (0)

#define PSG_CHECK_LAUNCH(name)                                                    \
  do {                                                                            \
    cudaError_t e__ = cudaGetLastError();                                         \
    if (e__ != cudaSuccess) {                                                     \
      psg_set_error("%s: launch failed: %s", name, cudaGetErrorString(e__));      \
      return PSG_ERR_CUDA;                                                        \
    }                                                                             \
    ++g_psg_launch_count;                                                         \
  } while (0)

#define PSG_DTYPE_F32 0
#define PSG_DTYPE_BF16 1

static inline int psg_num_sms() {
  static int sms = 0;
  if (sms == 0) {
    int dev = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    if (sms <= 0) sms = 148;
  }
  return sms;
}

// ---------------------------------------------------------------------------------------------
// device helpers
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ float psg_ld(const float* p) { return *p; }
__device__ __forceinline__ float psg_ld(const __nv_bfloat16* p) { return __bfloat162float(*p); }
__device__ __forceinline__ void psg_st(float* p, float v) { *p = v; }
__device__ __forceinline__ void psg_st(__nv_bfloat16* p, float v) { *p = __float2bfloat16_rn(v); }

template <typename T> struct Vec8;  // 8 consecutive elements of T, loaded/stored as vectors
template <> struct Vec8<float> {
  float v[8];
  __device__ __forceinline__ void load(const float* p) {
    float4 a = *reinterpret_cast<const float4*>(p);
    float4 b = *reinterpret_cast<const float4*>(p + 4);
    v[0] = a.x; v[1] = a.y; v[2] = a.z; v[3] = a.w; v[4] = b.x; v[5] = b.y; v[6] = b.z; v[7] = b.w;
  }
  __device__ __forceinline__ void store(float* p) const {
    *reinterpret_cast<float4*>(p) = make_float4(v[0], v[1], v[2], v[3]);
    *reinterpret_cast<float4*>(p + 4) = make_float4(v[4], v[5], v[6], v[7]);
  }
};
template <> struct Vec8<__nv_bfloat16> {
  float v[8];
  __device__ __forceinline__ void load(const __nv_bfloat16* p) {
    uint4 r = *reinterpret_cast<const uint4*>(p);
    const __nv_bfloat162* h = reinterpret_cast<const __nv_bfloat162*>(&r);
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      float2 f = __bfloat1622float2(h[i]);
      v[2 * i] = f.x; v[2 * i + 1] = f.y;
    }
  }
  __device__ __forceinline__ void store(__nv_bfloat16* p) const {
    uint4 r;
    __nv_bfloat162* h = reinterpret_cast<__nv_bfloat162*>(&r);
#pragma unroll
    for (int i = 0; i < 4; ++i) h[i] = __floats2bfloat162_rn(v[2 * i], v[2 * i + 1]);
    *reinterpret_cast<uint4*>(p) = r;
  }
};

__device__ __forceinline__ float psg_warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

__device__ __forceinline__ float psg_silu(float x) { return x / (1.f + __expf(-x)); }
__device__ __forceinline__ float psg_silu_grad(float x) {
  float s = 1.f / (1.f + __expf(-x));
  return s * (1.f + x * (1.f - s));
}
// exact (erf) GELU, as nn.GELU() default (reference: src/models/unet.py:183)
__device__ __forceinline__ float psg_gelu(float x) { return 0.5f * x * (1.f + erff(x * 0.70710678118654752f)); }
__device__ __forceinline__ float psg_gelu_grad(float x) {
  float cdf = 0.5f * (1.f + erff(x * 0.70710678118654752f));
  float pdf = 0.3989422804014327f * __expf(-0.5f * x * x);
  return cdf + x * pdf;
}

__device__ __forceinline__ float psg_rcp_approx(float x) { float r; asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x)); return r; }
__device__ __forceinline__ float psg_ex2_approx(float x) { float r; asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x)); return r; }

// Fast forms for the bf16 tensor-core epilogues (|error| < 2e-7, far below bf16 rounding): Abramowitz-Stegun 7.1.26
// for the normal CDF, evaluated on the tail side so that there is no cancellation; two MUFU ops (rcp, ex2) per value.
__device__ __forceinline__ void psg_gelu_parts(float x, float& cdf, float& e) {
  // z = |x|/sqrt(2); t = 1/(1 + p z); tail = 1 - Phi(|x|) = (poly(t)/2) * exp(-z^2), constants folded
  const float t = psg_rcp_approx(fmaf(0.23164189f, fabsf(x), 1.f));
  e = psg_ex2_approx(x * x * -0.72134752f);                       // exp(-x^2/2)
  const float poly = t * fmaf(t, fmaf(t, fmaf(t, fmaf(t, 0.5307027145f, -0.7265760135f), 0.7107068705f), -0.142248368f), 0.127414796f);
  const float tail = poly * e;
  cdf = x < 0.f ? tail : 1.f - tail;
}
__device__ __forceinline__ float psg_gelu_fast(float x) {
  float cdf, e;
  psg_gelu_parts(x, cdf, e);
  return x * cdf;
}
__device__ __forceinline__ float psg_gelu_grad_fast(float x) {
  float cdf, e;
  psg_gelu_parts(x, cdf, e);
  return fmaf(x * 0.3989422804014327f, e, cdf);
}
__device__ __forceinline__ float psg_silu_fast(float x) { return x * psg_rcp_approx(1.f + psg_ex2_approx(-1.4426950408889634f * x)); }
__device__ __forceinline__ float psg_silu_grad_fast(float x) {
  const float s = psg_rcp_approx(1.f + psg_ex2_approx(-1.4426950408889634f * x));
  return s * fmaf(x, 1.f - s, 1.f);
}

// Stateless counter-based dropout mask: keep iff hash(seed, idx) >= threshold (threshold = p * 2^32).  32-bit mixer
// (two multiply/xorshift rounds): it runs once per element inside GEMM epilogues, where a 64-bit mixer was the bottleneck.
__device__ __forceinline__ uint32_t psg_hash32(uint64_t seed, uint64_t idx) {
  uint32_t x = (uint32_t)idx ^ ((uint32_t)(idx >> 32) * 0x9E3779B9u) ^ (uint32_t)seed ^ ((uint32_t)(seed >> 32) * 0x85EBCA6Bu);
  x ^= x >> 16;
  x *= 0x7FEB352Du;
  x ^= x >> 15;
  x *= 0x846CA68Bu;
  x ^= x >> 16;
  return x;
}
// One hash decides two neighbouring elements (16 bits each): keep element idx iff its half of hash(seed, idx / 2) is
// >= threshold / 65536.  Every dropout site (GEMM epilogues, softmax, dropout_scale) and its backward use this one rule.
__device__ __forceinline__ bool psg_drop_keep2(uint32_t pair_hash, int odd, uint32_t threshold) {
  const uint32_t field = odd ? (pair_hash >> 16) : (pair_hash & 0xFFFFu);
  return field >= (threshold >> 16);
}
__device__ __forceinline__ bool psg_drop_keep(uint64_t seed, uint64_t idx, uint32_t threshold) {
  return psg_drop_keep2(psg_hash32(seed, idx >> 1), (int)(idx & 1), threshold);
}
