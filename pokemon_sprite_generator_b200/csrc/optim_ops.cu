// Optimiser-side bandwidth kernels over flat fp32 buffers: global L2 norm (+ finite flag), clip coefficient,
// fused AdamW.  SURVEY.md §2.1 K17/K18.
//
// Reference semantics replaced (src/training/improved_diffusion_trainer.py:399-413):
//   the 478 x `p.grad.norm(2).item()` loop + torch.nn.utils.clip_grad_norm_(max_norm) + torch.optim.AdamW(eps=1e-6).step()
// The clip coefficient and the "non-finite gradient -> skip the step" predicate stay on the device, so a
// training step needs no host synchronisation.
#include "psg_common.cuh"

namespace {

constexpr int kThreads = 256;
constexpr int kMaxBlocks = 1024;

// workspace layout: [kMaxBlocks] float partial sums | uint32 counter | (pad)
__global__ void __launch_bounds__(kThreads)
sumsq_kernel(const float* __restrict__ x, size_t n, float* __restrict__ partials, unsigned int* __restrict__ counter,
             float* __restrict__ out_sumsq, int accumulate) {
  float acc = 0.f;
  const size_t n4 = n >> 2;
  const float4* x4 = reinterpret_cast<const float4*>(x);
  for (size_t i = (size_t)blockIdx.x * kThreads + threadIdx.x; i < n4; i += (size_t)gridDim.x * kThreads) {
    const float4 v = x4[i];
    acc += v.x * v.x + v.y * v.y + v.z * v.z + v.w * v.w;
  }
  for (size_t i = (n4 << 2) + (size_t)blockIdx.x * kThreads + threadIdx.x; i < n; i += (size_t)gridDim.x * kThreads) acc += x[i] * x[i];
  __shared__ float warp_acc[kThreads / 32];
  __shared__ bool is_last;
  acc = psg_warp_sum(acc);
  if ((threadIdx.x & 31) == 0) warp_acc[threadIdx.x >> 5] = acc;
  __syncthreads();
  if (threadIdx.x == 0) {
    float s = 0.f;
    for (int w = 0; w < kThreads / 32; ++w) s += warp_acc[w];
    partials[blockIdx.x] = s;
    __threadfence();
    is_last = (atomicAdd(counter, 1u) == gridDim.x - 1);
  }
  __syncthreads();
  if (is_last) {
    __threadfence();
    double s = 0.0;
    for (int i = threadIdx.x; i < (int)gridDim.x; i += kThreads) s += (double)__ldcg(partials + i);
    __shared__ double dsh[kThreads];
    dsh[threadIdx.x] = s;
    __syncthreads();
    for (int o = kThreads / 2; o > 0; o >>= 1) {
      if (threadIdx.x < o) dsh[threadIdx.x] += dsh[threadIdx.x + o];
      __syncthreads();
    }
    if (threadIdx.x == 0) {
      *out_sumsq = accumulate ? *out_sumsq + (float)dsh[0] : (float)dsh[0];
      *counter = 0u;
    }
  }
}

// state[0] = total_norm, state[1] = clip coefficient (<= 1), state[2] = 1.0 if finite else 0.0,
// state[3] (when count_steps) = number of APPLIED optimiser steps so far: incremented only for finite gradients, so that
// Adam's bias correction does not advance over skipped batches (the reference `continue`s before optimizer.step(),
// src/training/improved_diffusion_trainer.py:395-397)
__global__ void clip_coef_kernel(const float* __restrict__ sumsq, float max_norm, float* __restrict__ state, int count_steps) {
  const float norm = sqrtf(*sumsq);
  const bool finite = isfinite(norm);
  float coef = 1.f;
  if (max_norm > 0.f) coef = fminf(max_norm / (norm + 1e-6f), 1.f);   // torch.nn.utils.clip_grad_norm_
  state[0] = norm;
  state[1] = finite ? coef : 0.f;
  state[2] = finite ? 1.f : 0.f;
  if (count_steps && finite) state[3] += 1.f;
}

__global__ void __launch_bounds__(kThreads)
adamw_kernel(float* __restrict__ p, const float* __restrict__ g, float* __restrict__ m, float* __restrict__ v, size_t n,
             float lr, float beta1, float beta2, float omb1, float omb2, double beta1_d, double beta2_d, float eps, float weight_decay,
             float bc1, float bc2_sqrt, const float* __restrict__ state, __nv_bfloat16* __restrict__ shadow, int coupled_l2, int step_from_state) {
  float gscale = 1.f;
  if (state != nullptr) {
    if (state[2] == 0.f) return;  // non-finite gradients: skip the whole step (the bf16 shadow stays valid)
    gscale = state[1];
    if (step_from_state) {        // bias corrections from the device-side count of applied steps (state[3] >= 1 here)
      const double t = (double)state[3];
      bc1 = (float)(1.0 - pow(beta1_d, t));
      bc2_sqrt = (float)sqrt(1.0 - pow(beta2_d, t));
    }
  }
  const float step = lr / bc1;
  // decoupled (torch.optim.AdamW): p *= 1 - lr*wd;  coupled (torch.optim.Adam(weight_decay=wd)): g += wd * p
  const float decay = coupled_l2 ? 1.f : 1.f - lr * weight_decay;
  const float l2 = coupled_l2 ? weight_decay : 0.f;
  const size_t n4 = n >> 2;
  float4* p4 = reinterpret_cast<float4*>(p);
  const float4* g4 = reinterpret_cast<const float4*>(g);
  float4* m4 = reinterpret_cast<float4*>(m);
  float4* v4 = reinterpret_cast<float4*>(v);
  auto upd = [&](float& pp, float gg, float& mm, float& vv) {
    gg = gg * gscale + l2 * pp;
    pp *= decay;
    mm = beta1 * mm + omb1 * gg;          // 1 - beta formed in double on the host, as torch does (1.f - 0.999f is off by 5e-5)
    vv = beta2 * vv + omb2 * gg * gg;
    const float denom = sqrtf(vv) / bc2_sqrt + eps;
    pp -= step * (mm / denom);
  };
  for (size_t i = (size_t)blockIdx.x * kThreads + threadIdx.x; i < n4; i += (size_t)gridDim.x * kThreads) {
    float4 pp = p4[i], gg = g4[i], mm = m4[i], vv = v4[i];
    upd(pp.x, gg.x, mm.x, vv.x); upd(pp.y, gg.y, mm.y, vv.y); upd(pp.z, gg.z, mm.z, vv.z); upd(pp.w, gg.w, mm.w, vv.w);
    p4[i] = pp; m4[i] = mm; v4[i] = vv;
    if (shadow) {   // bf16 copy of the updated parameters in the same layout: what the tensor-core GEMMs read
      __nv_bfloat162 lo = __floats2bfloat162_rn(pp.x, pp.y), hi = __floats2bfloat162_rn(pp.z, pp.w);
      uint2 pk;
      pk.x = *reinterpret_cast<uint32_t*>(&lo);
      pk.y = *reinterpret_cast<uint32_t*>(&hi);
      reinterpret_cast<uint2*>(shadow)[i] = pk;
    }
  }
  for (size_t i = (n4 << 2) + (size_t)blockIdx.x * kThreads + threadIdx.x; i < n; i += (size_t)gridDim.x * kThreads)
  {
    upd(p[i], g[i], m[i], v[i]);
    if (shadow) shadow[i] = __float2bfloat16_rn(p[i]);
  }
}

__global__ void __launch_bounds__(kThreads) cast_bf16_kernel(const float* __restrict__ x, __nv_bfloat16* __restrict__ y, size_t n) {
  const size_t n4 = n >> 2;
  const float4* x4 = reinterpret_cast<const float4*>(x);
  for (size_t i = (size_t)blockIdx.x * kThreads + threadIdx.x; i < n4; i += (size_t)gridDim.x * kThreads) {
    const float4 v = x4[i];
    __nv_bfloat162 lo = __floats2bfloat162_rn(v.x, v.y), hi = __floats2bfloat162_rn(v.z, v.w);
    uint2 pk;
    pk.x = *reinterpret_cast<uint32_t*>(&lo);
    pk.y = *reinterpret_cast<uint32_t*>(&hi);
    reinterpret_cast<uint2*>(y)[i] = pk;
  }
  for (size_t i = (n4 << 2) + (size_t)blockIdx.x * kThreads + threadIdx.x; i < n; i += (size_t)gridDim.x * kThreads)
    y[i] = __float2bfloat16_rn(x[i]);
}

__global__ void __launch_bounds__(kThreads) scale_kernel(float* __restrict__ x, size_t n, const float* __restrict__ state, float extra) {
  const float s = (state ? state[1] : 1.f) * extra;
  for (size_t i = (size_t)blockIdx.x * kThreads + threadIdx.x; i < n; i += (size_t)gridDim.x * kThreads) x[i] *= s;
}

}  // namespace

extern "C" {

// out_sumsq (=|+=) sum(x^2).  workspace: >= 1024 floats + 8 bytes, zeroed once.
int psg_sumsq(const float* x, long long n, float* out_sumsq, int accumulate, void* workspace, void* stream) {
  PSG_CHECK_ARG(x && out_sumsq && workspace && n > 0, "psg_sumsq: bad args");
  PSG_CHECK_ARG((uintptr_t)x % 16 == 0, "psg_sumsq: x must be 16B aligned");
  float* partials = (float*)workspace;
  unsigned int* counter = (unsigned int*)(partials + kMaxBlocks);
  long long g = (n / 4 + kThreads - 1) / kThreads;
  int cap = psg_num_sms() * 4;
  if (cap > kMaxBlocks) cap = kMaxBlocks;
  if (g > cap) g = cap;
  if (g < 1) g = 1;
  sumsq_kernel<<<(int)g, kThreads, 0, (cudaStream_t)stream>>>(x, (size_t)n, partials, counter, out_sumsq, accumulate);
  PSG_CHECK_LAUNCH("psg_sumsq");
  return PSG_OK;
}

// state (3 floats on device): total_norm, clip coefficient, finite flag.
int psg_clip_coef(const float* sumsq, float max_norm, float* state, void* stream) {
  PSG_CHECK_ARG(sumsq && state, "psg_clip_coef: null pointer");
  clip_coef_kernel<<<1, 1, 0, (cudaStream_t)stream>>>(sumsq, max_norm, state, 0);
  PSG_CHECK_LAUNCH("psg_clip_coef");
  return PSG_OK;
}

// Same, and state[3] (a 4th float) counts the steps that will actually be applied (finite gradients only).
int psg_clip_coef_count(const float* sumsq, float max_norm, float* state4, void* stream) {
  PSG_CHECK_ARG(sumsq && state4, "psg_clip_coef_count: null pointer");
  clip_coef_kernel<<<1, 1, 0, (cudaStream_t)stream>>>(sumsq, max_norm, state4, 1);
  PSG_CHECK_LAUNCH("psg_clip_coef_count");
  return PSG_OK;
}

// torch.optim.AdamW / torch.optim.Adam (no amsgrad) over flat buffers.  `step` is the 1-based step count (bias corrections
// computed on the host in double, as PyTorch does); step <= 0: take the count of applied steps from state[3] (see
// psg_clip_coef_count).  coupled_l2 = 0: decoupled decay (AdamW, improved_diffusion_trainer.py:277-283);
// 1: L2 added to the gradient (the reference's `else` branch, torch.optim.Adam(weight_decay=...), :285-292).
// state may be null (no clipping / skipping).
int psg_adam_step(float* p, const float* g, float* m, float* v, long long n, float lr, double beta1, double beta2, float eps,
                  float weight_decay, long long step, int coupled_l2, const float* state, void* bf16_shadow, void* stream) {
  PSG_CHECK_ARG(p && g && m && v && n > 0, "psg_adam_step: bad args");
  PSG_CHECK_ARG(step >= 1 || state != nullptr, "psg_adam_step: step <= 0 needs the device state (applied-step counter)");
  PSG_CHECK_ARG(((uintptr_t)p % 16 == 0) && ((uintptr_t)g % 16 == 0) && ((uintptr_t)m % 16 == 0) && ((uintptr_t)v % 16 == 0),
                "psg_adam_step: buffers must be 16B aligned");
  PSG_CHECK_ARG(bf16_shadow == nullptr || (uintptr_t)bf16_shadow % 8 == 0, "psg_adam_step: shadow must be 8B aligned");
  const double st = step >= 1 ? (double)step : 1.0;
  const double bc1 = 1.0 - pow(beta1, st);
  const double bc2 = 1.0 - pow(beta2, st);
  long long g_ = (n / 4 + kThreads - 1) / kThreads;
  long long cap = (long long)psg_num_sms() * 16;
  if (g_ > cap) g_ = cap;
  if (g_ < 1) g_ = 1;
  adamw_kernel<<<(int)g_, kThreads, 0, (cudaStream_t)stream>>>(p, g, m, v, (size_t)n, lr, (float)beta1, (float)beta2, (float)(1.0 - beta1), (float)(1.0 - beta2), beta1, beta2, eps, weight_decay, (float)bc1,
                                                              (float)sqrt(bc2), state, (__nv_bfloat16*)bf16_shadow, coupled_l2 ? 1 : 0,
                                                              step >= 1 ? 0 : 1);
  PSG_CHECK_LAUNCH("psg_adam_step");
  return PSG_OK;
}

int psg_adamw_step(float* p, const float* g, float* m, float* v, long long n, float lr, double beta1, double beta2, float eps,
                   float weight_decay, long long step, const float* state, void* bf16_shadow, void* stream) {
  PSG_CHECK_ARG(step >= 1, "psg_adamw_step: step must be >= 1");
  return psg_adam_step(p, g, m, v, n, lr, beta1, beta2, eps, weight_decay, step, 0, state, bf16_shadow, stream);
}

// y = bf16(x): refreshes the bf16 shadow of the flat parameter buffer after the parameters changed outside psg_adamw_step.
int psg_cast_bf16(const float* x, void* y, long long n, void* stream) {
  PSG_CHECK_ARG(x && y && n > 0, "psg_cast_bf16: bad args");
  PSG_CHECK_ARG(((uintptr_t)x % 16 == 0) && ((uintptr_t)y % 8 == 0), "psg_cast_bf16: buffers must be 16B / 8B aligned");
  long long g_ = (n / 4 + kThreads - 1) / kThreads;
  long long cap = (long long)psg_num_sms() * 16;
  if (g_ > cap) g_ = cap;
  if (g_ < 1) g_ = 1;
  cast_bf16_kernel<<<(int)g_, kThreads, 0, (cudaStream_t)stream>>>(x, (__nv_bfloat16*)y, (size_t)n);
  PSG_CHECK_LAUNCH("psg_cast_bf16");
  return PSG_OK;
}

// x *= state[1] * extra  (apply the clip coefficient in place, e.g. for an external optimiser; extra = 1/world_size)
int psg_scale_inplace(float* x, long long n, const float* state, float extra, void* stream) {
  PSG_CHECK_ARG(x && n > 0, "psg_scale_inplace: bad args");
  long long g_ = (n + kThreads - 1) / kThreads;
  long long cap = (long long)psg_num_sms() * 16;
  if (g_ > cap) g_ = cap;
  scale_kernel<<<(int)g_, kThreads, 0, (cudaStream_t)stream>>>(x, (size_t)n, state, extra);
  PSG_CHECK_LAUNCH("psg_scale_inplace");
  return PSG_OK;
}

}  // extern "C"
