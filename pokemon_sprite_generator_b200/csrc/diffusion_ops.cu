// Bandwidth-bound diffusion kernels: q_sample (K14), SmoothL1 fwd+bwd (K15), DDPM reverse step (K16).
//
// Reference semantics:
//   q_sample    : NoiseScheduler.add_noise   src/training/improved_diffusion_trainer.py:50-65
//   smooth_l1   : nn.SmoothL1Loss(beta=0.1)  src/training/improved_diffusion_trainer.py:300,388
//   ddpm_step 0 : ddpm_sample update         src/training/improved_diffusion_trainer.py:543-567
//   ddpm_step 1 : sample_previous_timestep   src/training/final_trainer.py:52-71
//   reverse_step 2 : sample_prev_timestep    src/training/diffusers_trainer.py:76-100  (x0-prediction form)
//   reverse_step 3 : gradio sampling loop    gradio_app.py:324-361                     (denoise, then re-noise to the next step)
//
// Bit-exactness: eager PyTorch evaluates each tensor op separately (no FMA contraction), so the
// arithmetic here is spelled with __fmul_rn/__fadd_rn/__fsub_rn/__fdiv_rn in the reference's order.
#include "psg_common.cuh"

namespace {

constexpr int kThreads = 256;

__device__ __forceinline__ float clampf(float v, float lo, float hi) { return fminf(fmaxf(v, lo), hi); }

// one CTA row per (sample, chunk); float4 path
__global__ void __launch_bounds__(kThreads)
q_sample_kernel(const float* __restrict__ x0, const float* __restrict__ noise, const long long* __restrict__ t,
                const float* __restrict__ sqrt_ac, const float* __restrict__ sqrt_1mac, float* __restrict__ out,
                int n_per, int num_t, int do_clamp, float clo, float chi, int* __restrict__ nonfinite, int vec_ok) {
  const int b = blockIdx.y;
  long long ti = t[b];
  ti = ti < 0 ? 0 : (ti >= num_t ? num_t - 1 : ti);
  const float a = sqrt_ac[ti];
  const float s = sqrt_1mac[ti];
  const size_t base = (size_t)b * n_per;
  int bad = 0;
  if (vec_ok) {             // n_per % 4 == 0 and 16-byte aligned bases (decided by the launcher)
    const int n4 = n_per >> 2;
    const float4* x4 = reinterpret_cast<const float4*>(x0 + base);
    const float4* e4 = reinterpret_cast<const float4*>(noise + base);
    float4* o4 = reinterpret_cast<float4*>(out + base);
    for (int i = blockIdx.x * kThreads + threadIdx.x; i < n4; i += gridDim.x * kThreads) {
      float4 x = x4[i], e = e4[i], o;
      if (do_clamp) { x.x = clampf(x.x, clo, chi); x.y = clampf(x.y, clo, chi); x.z = clampf(x.z, clo, chi); x.w = clampf(x.w, clo, chi); }
      o.x = __fadd_rn(__fmul_rn(a, x.x), __fmul_rn(s, e.x));
      o.y = __fadd_rn(__fmul_rn(a, x.y), __fmul_rn(s, e.y));
      o.z = __fadd_rn(__fmul_rn(a, x.z), __fmul_rn(s, e.z));
      o.w = __fadd_rn(__fmul_rn(a, x.w), __fmul_rn(s, e.w));
      bad |= !(isfinite(o.x) && isfinite(o.y) && isfinite(o.z) && isfinite(o.w));
      o4[i] = o;
    }
  } else {
    for (int i = blockIdx.x * kThreads + threadIdx.x; i < n_per; i += gridDim.x * kThreads) {
      float x = x0[base + i];
      if (do_clamp) x = clampf(x, clo, chi);
      float o = __fadd_rn(__fmul_rn(a, x), __fmul_rn(s, noise[base + i]));
      bad |= !isfinite(o);
      out[base + i] = o;
    }
  }
  if (nonfinite != nullptr && __syncthreads_or(bad) && threadIdx.x == 0) atomicOr(nonfinite, 1);
}

// reference fallback (improved_diffusion_trainer.py:61-63): if any element was NaN/Inf, out = x0 + 0.1*noise
__global__ void __launch_bounds__(kThreads)
q_sample_fallback_kernel(const float* __restrict__ x0, const float* __restrict__ noise, float* __restrict__ out,
                         size_t n, int do_clamp, float clo, float chi, const int* __restrict__ nonfinite) {
  if (*nonfinite == 0) return;
  for (size_t i = (size_t)blockIdx.x * kThreads + threadIdx.x; i < n; i += (size_t)gridDim.x * kThreads) {
    float x = x0[i];
    if (do_clamp) x = clampf(x, clo, chi);
    out[i] = __fadd_rn(x, __fmul_rn(0.1f, noise[i]));
  }
}

// SmoothL1: per-block partial sums, last block folds them in a fixed order (deterministic).
__global__ void __launch_bounds__(kThreads)
smooth_l1_kernel(const float* __restrict__ pred, const float* __restrict__ target, float* __restrict__ dpred,
                 float* __restrict__ loss, float* __restrict__ partials, unsigned int* __restrict__ counter,
                 size_t n, float beta, float grad_scale) {
  float acc = 0.f;
  const float inv_beta = 1.f / beta;
  const float gs = grad_scale / (float)n;
  const size_t n4 = n >> 2;
  const float4* p4 = reinterpret_cast<const float4*>(pred);
  const float4* t4 = reinterpret_cast<const float4*>(target);
  float4* g4 = reinterpret_cast<float4*>(dpred);
  for (size_t i = (size_t)blockIdx.x * kThreads + threadIdx.x; i < n4; i += (size_t)gridDim.x * kThreads) {
    float4 p = p4[i], t = t4[i], g;
    float d[4] = {p.x - t.x, p.y - t.y, p.z - t.z, p.w - t.w};
    float gg[4];
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      float ad = fabsf(d[k]);
      if (ad < beta) { acc += 0.5f * d[k] * d[k] * inv_beta; gg[k] = d[k] * inv_beta * gs; }
      else           { acc += ad - 0.5f * beta;              gg[k] = (d[k] > 0.f ? gs : (d[k] < 0.f ? -gs : 0.f)); }
    }
    if (dpred != nullptr) { g.x = gg[0]; g.y = gg[1]; g.z = gg[2]; g.w = gg[3]; g4[i] = g; }
  }
  // tail
  for (size_t i = (n4 << 2) + (size_t)blockIdx.x * kThreads + threadIdx.x; i < n; i += (size_t)gridDim.x * kThreads) {
    float d = pred[i] - target[i];
    float ad = fabsf(d), g;
    if (ad < beta) { acc += 0.5f * d * d * inv_beta; g = d * inv_beta * gs; }
    else           { acc += ad - 0.5f * beta;        g = (d > 0.f ? gs : (d < 0.f ? -gs : 0.f)); }
    if (dpred != nullptr) dpred[i] = g;
  }
  __shared__ float warp_acc[kThreads / 32];
  __shared__ bool is_last;
  acc = psg_warp_sum(acc);
  if ((threadIdx.x & 31) == 0) warp_acc[threadIdx.x >> 5] = acc;
  __syncthreads();
  if (threadIdx.x == 0) {
    float s = 0.f;
#pragma unroll
    for (int w = 0; w < kThreads / 32; ++w) s += warp_acc[w];
    partials[blockIdx.x] = s;
    __threadfence();
    unsigned int done = atomicAdd(counter, 1u);
    is_last = (done == gridDim.x - 1);
  }
  __syncthreads();
  if (is_last) {
    __threadfence();
    float s = 0.f;
    for (int i = threadIdx.x; i < (int)gridDim.x; i += kThreads) s += __ldcg(partials + i);
    s = psg_warp_sum(s);
    if ((threadIdx.x & 31) == 0) warp_acc[threadIdx.x >> 5] = s;
    __syncthreads();
    if (threadIdx.x == 0) {
      float tot = 0.f;
#pragma unroll
      for (int w = 0; w < kThreads / 32; ++w) tot += warp_acc[w];
      *loss = tot / (float)n;
      *counter = 0u;  // re-arm for the next call
    }
  }
}

// DDPM reverse step.  tables live on the device; t is a host scalar (same t for the whole batch, as in
// both reference samplers).
__global__ void __launch_bounds__(kThreads)
ddpm_step_kernel(const float* __restrict__ x, const float* __restrict__ eps, const float* __restrict__ z,
                 float* __restrict__ out, size_t n, int mode, const float* __restrict__ tab0,
                 const float* __restrict__ tab1, const float* __restrict__ tab2, const float* __restrict__ tab3, int t) {
  // mode 0: tab0 = 1/sqrt(alpha), tab1 = beta/sqrt(1-abar), tab2 = sqrt(beta)
  // mode 1: tab0 = sqrt(1/alpha), tab1 = beta, tab2 = sqrt(1-abar), tab3 = sqrt(posterior_variance)
  const float c0 = tab0[t], c1 = tab1[t], c2 = tab2[t];
  const float sig = (mode == 0) ? c2 : (z != nullptr ? tab3[t] : 0.f);
  const size_t n4 = n >> 2;
  const float4* x4 = reinterpret_cast<const float4*>(x);
  const float4* e4 = reinterpret_cast<const float4*>(eps);
  const float4* z4 = reinterpret_cast<const float4*>(z);
  float4* o4 = reinterpret_cast<float4*>(out);
  auto upd = [&](float xv, float ev, float zv) -> float {
    float m;
    if (mode == 0) m = __fmul_rn(c0, __fsub_rn(xv, __fmul_rn(c1, ev)));
    else           m = __fmul_rn(c0, __fsub_rn(xv, __fdiv_rn(__fmul_rn(c1, ev), c2)));
    if (z != nullptr) m = __fadd_rn(m, __fmul_rn(sig, zv));
    return m;
  };
  for (size_t i = (size_t)blockIdx.x * kThreads + threadIdx.x; i < n4; i += (size_t)gridDim.x * kThreads) {
    float4 xv = x4[i], ev = e4[i], zv = make_float4(0.f, 0.f, 0.f, 0.f), o;
    if (z != nullptr) zv = z4[i];
    o.x = upd(xv.x, ev.x, zv.x); o.y = upd(xv.y, ev.y, zv.y); o.z = upd(xv.z, ev.z, zv.z); o.w = upd(xv.w, ev.w, zv.w);
    o4[i] = o;
  }
  for (size_t i = (n4 << 2) + (size_t)blockIdx.x * kThreads + threadIdx.x; i < n; i += (size_t)gridDim.x * kThreads)
    out[i] = upd(x[i], eps[i], z != nullptr ? z[i] : 0.f);
}

inline int grid_for(size_t work_items, int max_blocks) {
  size_t g = (work_items + kThreads - 1) / kThreads;
  if (g < 1) g = 1;
  if (g > (size_t)max_blocks) g = max_blocks;
  return (int)g;
}

}  // namespace

extern "C" {

// Replaces NoiseScheduler.add_noise (improved_diffusion_trainer.py:50-65) and the preceding
// clamp(latent, -3, 3) (:363) when do_clamp != 0.  nonfinite (device int, may be null) is OR-ed with 1 when
// any output is NaN/Inf; when it is given the reference's `x0 + 0.1*noise` fallback is applied on device.
int psg_q_sample(const float* x0, const float* noise, const long long* t, const float* sqrt_ac,
                 const float* sqrt_1mac, float* out, int batch, int n_per, int num_t, int do_clamp,
                 float clamp_lo, float clamp_hi, int* nonfinite, void* stream) {
  PSG_CHECK_ARG(batch >= 0 && n_per >= 0 && num_t > 0, "psg_q_sample: bad sizes");
  if (batch == 0 || n_per == 0) return PSG_OK;  // empty batch: nothing to do (pointers may be null)
  PSG_CHECK_ARG(x0 && noise && t && sqrt_ac && sqrt_1mac && out, "psg_q_sample: null pointer");
  PSG_CHECK_ARG(batch <= 65535, "psg_q_sample: batch > 65535");
  cudaStream_t s = (cudaStream_t)stream;
  dim3 grid(grid_for((size_t)(n_per + 3) / 4, 64), batch);
  // float4 path only for whole vectors at 16-byte aligned addresses (a sliced / offset view takes the scalar path)
  const int vec_ok = (n_per % 4 == 0) && ((uintptr_t)x0 % 16 == 0) && ((uintptr_t)noise % 16 == 0) && ((uintptr_t)out % 16 == 0);
  q_sample_kernel<<<grid, kThreads, 0, s>>>(x0, noise, t, sqrt_ac, sqrt_1mac, out, n_per, num_t, do_clamp,
                                            clamp_lo, clamp_hi, nonfinite, vec_ok);
  PSG_CHECK_LAUNCH("psg_q_sample");
  if (nonfinite != nullptr) {
    size_t n = (size_t)batch * n_per;
    q_sample_fallback_kernel<<<grid_for(n, psg_num_sms() * 8), kThreads, 0, s>>>(x0, noise, out, n, do_clamp,
                                                                                   clamp_lo, clamp_hi, nonfinite);
    PSG_CHECK_LAUNCH("psg_q_sample(fallback)");
  }
  return PSG_OK;
}

// workspace: 1024 floats of partial sums followed by one uint32 counter (zero on first use): >= 4100 bytes.
int psg_smooth_l1_fwd_bwd(const float* pred, const float* target, float* dpred, float* loss, void* workspace,
                          long long n, float beta, float grad_scale, void* stream) {
  PSG_CHECK_ARG(pred && target && loss && workspace, "psg_smooth_l1_fwd_bwd: null pointer");
  PSG_CHECK_ARG(n > 0 && beta > 0.f, "psg_smooth_l1_fwd_bwd: bad n/beta");
  PSG_CHECK_ARG(((uintptr_t)pred % 16 == 0) && ((uintptr_t)target % 16 == 0) && ((uintptr_t)dpred % 16 == 0),
                "psg_smooth_l1_fwd_bwd: pointers must be 16B aligned");
  float* partials = (float*)workspace;
  unsigned int* counter = (unsigned int*)(partials + 1024);
  int grid = grid_for((size_t)(n + 3) / 4, 1024);
  int cap = psg_num_sms() * 4;
  if (grid > cap) grid = cap;
  smooth_l1_kernel<<<grid, kThreads, 0, (cudaStream_t)stream>>>(pred, target, dpred, loss, partials, counter,
                                                                 (size_t)n, beta, grad_scale);
  PSG_CHECK_LAUNCH("psg_smooth_l1_fwd_bwd");
  return PSG_OK;
}

// The two remaining reverse-step variants of the reference, each elementwise operation rounded separately and in the reference's
// order (no FMA contraction), with the per-step scalars `c` evaluated by the caller exactly as the reference evaluates them:
//   mode 2: pred = (x - c0 * eps) / c1;  prev = c2 * pred + c3 * eps;  [+ c4 * z]
//           c = {sqrt(1-abar_t), sqrt(abar_t), sqrt(abar_prev), sqrt(1-abar_prev), sqrt(posterior_variance_t)}
//   mode 3: lat = (x - c0 * eps) / c1;   [lat = c2 * lat + c3 * z]
//           c = {(1-alpha_t)/sqrt(1-abar_t), sqrt(alpha_t), sqrt(alpha_next), sqrt(1-alpha_next)}
struct RevCoef { float c[5]; };
__global__ void __launch_bounds__(kThreads)
reverse_step_kernel(const float* __restrict__ x, const float* __restrict__ eps, const float* __restrict__ z, float* __restrict__ out,
                    size_t n, int mode, RevCoef k) {
  for (size_t i = (size_t)blockIdx.x * kThreads + threadIdx.x; i < n; i += (size_t)gridDim.x * kThreads) {
    const float xv = x[i], ev = eps[i];
    float r = __fdiv_rn(__fsub_rn(xv, __fmul_rn(k.c[0], ev)), k.c[1]);
    if (mode == 2) {
      r = __fadd_rn(__fmul_rn(k.c[2], r), __fmul_rn(k.c[3], ev));
      if (z != nullptr) r = __fadd_rn(r, __fmul_rn(k.c[4], z[i]));
    } else if (z != nullptr) {
      r = __fadd_rn(__fmul_rn(k.c[2], r), __fmul_rn(k.c[3], z[i]));
    }
    out[i] = r;
  }
}

__global__ void __launch_bounds__(kThreads)
reparam_kernel(const float* __restrict__ mu, const float* __restrict__ logvar, const float* __restrict__ eps, float* __restrict__ out,
               size_t n, int do_clamp, float lo, float hi) {
  for (size_t i = (size_t)blockIdx.x * kThreads + threadIdx.x; i < n; i += (size_t)gridDim.x * kThreads) {
    const float sd = expf(__fmul_rn(0.5f, logvar[i]));
    float v = __fadd_rn(mu[i], __fmul_rn(eps[i], sd));
    if (do_clamp) v = fminf(fmaxf(v, lo), hi);
    out[i] = v;
  }
}

int psg_reparam(const float* mu, const float* logvar, const float* eps, float* out, long long n, int do_clamp, float lo, float hi,
                void* stream) {
  if (n <= 0) return PSG_OK;
  PSG_CHECK_ARG(mu && logvar && eps && out, "psg_reparam: null pointer");
  reparam_kernel<<<grid_for((size_t)n, psg_num_sms() * 8), kThreads, 0, (cudaStream_t)stream>>>(mu, logvar, eps, out, (size_t)n, do_clamp, lo, hi);
  PSG_CHECK_LAUNCH("psg_reparam");
  return PSG_OK;
}

int psg_reverse_step(const float* x, const float* eps, const float* z, float* out, long long n, int mode, const float* coef5,
                     void* stream) {
  if (n <= 0) return PSG_OK;
  PSG_CHECK_ARG(x && eps && out && coef5, "psg_reverse_step: null pointer");
  PSG_CHECK_ARG(mode == 2 || mode == 3, "psg_reverse_step: mode must be 2 (diffusers_trainer) or 3 (gradio loop)");
  RevCoef k;
  for (int i = 0; i < 5; ++i) k.c[i] = coef5[i];      // host array
  int grid = grid_for((size_t)n, psg_num_sms() * 8);
  reverse_step_kernel<<<grid, kThreads, 0, (cudaStream_t)stream>>>(x, eps, z, out, (size_t)n, mode, k);
  PSG_CHECK_LAUNCH("psg_reverse_step");
  return PSG_OK;
}

int psg_ddpm_step(const float* x, const float* eps, const float* z, float* out, long long n, int mode,
                  const float* tab0, const float* tab1, const float* tab2, const float* tab3, int t, int num_t,
                  void* stream) {
  if (n <= 0) return PSG_OK;
  PSG_CHECK_ARG(x && eps && out && tab0 && tab1 && tab2, "psg_ddpm_step: null pointer");
  PSG_CHECK_ARG(mode == 0 || (mode == 1 && tab3), "psg_ddpm_step: bad mode/tables");
  PSG_CHECK_ARG(t >= 0 && t < num_t, "psg_ddpm_step: t=%d out of range [0,%d)", t, num_t);
  PSG_CHECK_ARG(((uintptr_t)x % 16 == 0) && ((uintptr_t)eps % 16 == 0) && ((uintptr_t)z % 16 == 0) &&
                    ((uintptr_t)out % 16 == 0), "psg_ddpm_step: pointers must be 16B aligned");
  if (n <= 0) return PSG_OK;
  int grid = grid_for((size_t)(n + 3) / 4, psg_num_sms() * 8);
  ddpm_step_kernel<<<grid, kThreads, 0, (cudaStream_t)stream>>>(x, eps, z, out, (size_t)n, mode, tab0, tab1, tab2,
                                                                 tab3, t);
  PSG_CHECK_LAUNCH("psg_ddpm_step");
  return PSG_OK;
}

}  // extern "C"
