// Fused GEMM/conv epilogue shared by the tcgen05 engine (umma_gemm.cu) and the SIMT fp32 engine
// (simt_gemm.cu).  One definition so that both engines produce the same function of the accumulator:
//
//   v   = acc + bias[n] + rowbias[(m / rows_per_group) * ld_rowbias + n]
//   (aux_out[m,n] = v)                       -- optional pre-activation store, activation dtype            [aux_act == NONE]
//   (aux_out[m,n] = act'(v) * dropmask * keep_scale)  -- or the saved local derivative of act+dropout     [aux_act != NONE]
//   v   = act(v)                             -- none | gelu(erf) | silu | relu (forward only)
//   v  *= act'(aux_in[m,n])                  -- optional, backward of a fused activation; aux_act == PSG_ACT_MUL: v *= aux_in
//   v   = dropout(v; seed, threshold) * keep_scale
//   v   = alpha * v + residual[m,n]          -- residual in activation dtype, may alias out
//   out[m,n] (+)= v                          -- fp32 or bf16; accumulate only for fp32
//
// This covers every fused op of SURVEY.md §2.1 K1/K3/K6-K9: conv bias + time/text broadcast add
// (reference src/models/unet.py:116-124), out-proj `x*0.7 + residual` (:220-221), `x*0.8 + residual`
// (:238-239), FFN GELU / `x*0.6 + residual` (:181-187,249-251).
#pragma once
#include "psg_common.cuh"

#define PSG_ACT_NONE 0
#define PSG_ACT_GELU 1
#define PSG_ACT_SILU 2
#define PSG_ACT_MUL 3      // aux_in only: multiply by the stored value itself (a derivative saved by the forward epilogue)
#define PSG_ACT_RELU 4     // forward only (the VAE encoder's stem, src/models/vae_decoder.py:77-88); no saved derivative

struct PsgEpilogue {
  void* out;               // [M, ldc]
  long long ldc;
  int out_dtype;           // PSG_DTYPE_*
  int act_dtype;           // dtype of residual / aux tensors
  const float* bias;       // [N] or null
  const float* rowbias;    // [groups, ld_rowbias] or null
  int rows_per_group;
  long long ld_rowbias;
  int act;                 // PSG_ACT_*
  float alpha;
  const void* residual;    // [M, ldr] or null
  long long ldr;
  void* aux_out;           // [M, ld_aux] or null
  const void* aux_in;      // [M, ld_aux] or null
  long long ld_aux;
  int aux_act;             // activation whose derivative multiplies (PSG_ACT_*), used with aux_in
  int accumulate;          // out += v (fp32 out only)
  unsigned long long drop_seed;
  unsigned int drop_threshold;  // 0 = no dropout; keep iff hash >= threshold
  float drop_scale;             // 1/(1-p)
};

__device__ __forceinline__ float psg_epi_load_act(const void* p, int dtype, long long idx) {
  return dtype == PSG_DTYPE_BF16 ? __bfloat162float(reinterpret_cast<const __nv_bfloat16*>(p)[idx])
                                 : reinterpret_cast<const float*>(p)[idx];
}
__device__ __forceinline__ void psg_epi_store(void* p, int dtype, long long idx, float v) {
  if (dtype == PSG_DTYPE_BF16) reinterpret_cast<__nv_bfloat16*>(p)[idx] = __float2bfloat16_rn(v);
  else reinterpret_cast<float*>(p)[idx] = v;
}

// Scalar form (SIMT engine, tails).
__device__ __forceinline__ void psg_epilogue_scalar(const PsgEpilogue& e, float acc, long long m, long long n, long long N) {
  float v = acc;
  if (e.bias) v += e.bias[n];
  if (e.rowbias) v += e.rowbias[(m / e.rows_per_group) * e.ld_rowbias + n];
  const float keep = (!e.drop_threshold || psg_drop_keep(e.drop_seed, (uint64_t)(m * N + n), e.drop_threshold)) ? e.drop_scale : 0.f;
  if (e.aux_out) {
    float s = v;
    if (e.aux_act != PSG_ACT_NONE)
      s = keep * ((e.act == PSG_ACT_GELU) ? psg_gelu_grad(v) : (e.act == PSG_ACT_SILU ? psg_silu_grad(v) : 1.f));
    psg_epi_store(e.aux_out, e.act_dtype, m * e.ld_aux + n, s);
  }
  if (e.act == PSG_ACT_GELU) v = psg_gelu(v);
  else if (e.act == PSG_ACT_SILU) v = psg_silu(v);
  else if (e.act == PSG_ACT_RELU) v = fmaxf(v, 0.f);
  if (e.aux_in) {
    float a = psg_epi_load_act(e.aux_in, e.act_dtype, m * e.ld_aux + n);
    v *= (e.aux_act == PSG_ACT_GELU) ? psg_gelu_grad(a) : (e.aux_act == PSG_ACT_SILU ? psg_silu_grad(a) : (e.aux_act == PSG_ACT_MUL ? a : 1.f));
  }
  if (e.drop_threshold) v *= keep;
  v *= e.alpha;
  if (e.residual) v += psg_epi_load_act(e.residual, e.act_dtype, m * e.ldr + n);
  long long o = m * e.ldc + n;
  if (e.out_dtype == PSG_DTYPE_BF16) {
    reinterpret_cast<__nv_bfloat16*>(e.out)[o] = __float2bfloat16_rn(v);
  } else {
    float* po = reinterpret_cast<float*>(e.out) + o;
    *po = e.accumulate ? (*po + v) : v;
  }
}
