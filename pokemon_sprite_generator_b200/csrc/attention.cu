// Fused multi-head attention core (softmax(QK^T/sqrt(d)) V), forward and backward, flash-style: the
// [B, h, Lq, Lk] score matrix is never materialised.  SURVEY.md §2.1 K6/K7.
//
// Reference semantics: nn.MultiheadAttention(batch_first=True) core inside CrossAttentionBlock
// (src/models/unet.py:160-173,217,235): scale 1/sqrt(head_dim), softmax over keys, dropout(p) on the
// probabilities in train mode, no masks.  Projections run in the GEMM engines; this file is only the core.
//
// v1 is a CUDA-core kernel (fp32 math, T = fp32 | bf16 storage): one CTA per (batch, head, query tile) keeps a
// K/V chunk in shared memory; one warp per query row runs an online softmax over key chunks.  Backward is two
// atomic-free passes: dQ (per query row) and dK/dV (per key row), both recomputing P from the saved
// log-sum-exp.  The attention core is 0.7% of the model FLOPs, so it is kept simple first.
#include "psg_common.cuh"
#include <math.h>

namespace attn {

constexpr int kWarps = 8;
constexpr int kThreads = kWarps * 32;
constexpr int kMaxHeadDim = 320;

template <typename T> struct WordTraits;
template <> struct WordTraits<float> {
  static constexpr int EPW = 1;  // elements per 32-bit word
  __device__ static void unpack(uint32_t w, float& a, float& b) { a = __uint_as_float(w); b = 0.f; }
};
template <> struct WordTraits<__nv_bfloat16> {
  static constexpr int EPW = 2;
  __device__ static void unpack(uint32_t w, float& a, float& b) {
    a = __uint_as_float(w << 16);
    b = __uint_as_float(w & 0xffff0000u);
  }
};

struct Params {
  const void *q, *k, *v;   // q: [B*Lq, ldq] (+ head*hd), k/v: [B*Lk, ldk]
  void* o;                 // [B*Lq, ldo]
  float* lse;              // [B, H, Lq]
  long long ldq, ldk, ldv, ldo;
  int B, H, Lq, Lk, hd;
  float scale;
  int kc;                  // keys (or queries in dkv pass) per smem chunk
  int pitch;               // smem row pitch in words (odd)
  unsigned long long drop_seed;
  unsigned int drop_threshold;
  float drop_scale;
  // backward
  const void* dout;        // [B*Lq, lddo]
  long long lddo;
  float* dsum;             // [B, H, Lq]  D_i = dO_i . O_i
  void *dq, *dk, *dv;
  long long lddq, lddk, lddv;
  const int* key_len;      // forward only, nullable: [B] number of valid keys per sample (a padding mask that is a suffix)
};

__device__ __forceinline__ bool keep(const Params& p, int b, int h, int i, int j) {
  if (p.drop_threshold == 0) return true;
  uint64_t idx = (((uint64_t)(b * p.H + h) * p.Lq + i) * p.Lk + j);
  return psg_drop_keep(p.drop_seed, idx, p.drop_threshold);
}

// copy rows [r0, r0+nr) x hd of a token-major matrix into smem words with row pitch `pitch`
template <typename T>
__device__ __forceinline__ void load_rows(uint32_t* dst, const T* src, long long ld, int r0, int nr, int hd, int pitch) {
  constexpr int EPW = WordTraits<T>::EPW;
  const int words = hd / EPW;
  for (int idx = threadIdx.x; idx < nr * words; idx += kThreads) {
    const int r = idx / words, w = idx - r * words;
    dst[r * pitch + w] = reinterpret_cast<const uint32_t*>(src + (long long)(r0 + r) * ld)[w];
  }
}

__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}

// dot of fp32 vector q[hd] (smem) with smem word row
template <typename T>
__device__ __forceinline__ float dot_row(const float* q, const uint32_t* row, int words) {
  float acc = 0.f;
  if (WordTraits<T>::EPW == 2) {
    for (int w = 0; w < words; ++w) {
      float a, b;
      WordTraits<T>::unpack(row[w], a, b);
      acc = fmaf(q[2 * w], a, acc);
      acc = fmaf(q[2 * w + 1], b, acc);
    }
  } else {
    for (int w = 0; w < words; ++w) acc = fmaf(q[w], __uint_as_float(row[w]), acc);
  }
  return acc;
}

// ---------------------------------------------------------------------------------------------------------
// forward
// ---------------------------------------------------------------------------------------------------------
template <typename T>
__global__ void __launch_bounds__(kThreads) attn_fwd_kernel(Params p) {
  constexpr int EPW = WordTraits<T>::EPW;
  constexpr int MAXW = kMaxHeadDim / EPW / 32;  // words per lane
  extern __shared__ uint32_t smem[];
  const int words = p.hd / EPW;
  uint32_t* sK = smem;
  uint32_t* sV = sK + p.kc * p.pitch;
  float* sQ = reinterpret_cast<float*>(sV + p.kc * p.pitch);  // [kWarps][hd]
  float* sP = sQ + kWarps * p.hd;                              // [kWarps][kc]
  const int bh = blockIdx.x, b = bh / p.H, h = bh - b * p.H;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int rows_per_cta = (p.Lq + gridDim.y - 1) / gridDim.y;
  const int i0 = blockIdx.y * rows_per_cta, i1 = min(p.Lq, i0 + rows_per_cta);
  const T* Q = reinterpret_cast<const T*>(p.q) + (long long)b * p.Lq * p.ldq + h * p.hd;
  const T* K = reinterpret_cast<const T*>(p.k) + (long long)b * p.Lk * p.ldk + h * p.hd;
  const T* V = reinterpret_cast<const T*>(p.v) + (long long)b * p.Lk * p.ldv + h * p.hd;
  T* O = reinterpret_cast<T*>(p.o) + (long long)b * p.Lq * p.ldo + h * p.hd;
  float* myQ = sQ + warp * p.hd;
  float* myP = sP + warp * p.kc;
  const int nrounds = (i1 - i0 + kWarps - 1) / kWarps;

  for (int rd = 0; rd < nrounds; ++rd) {
    const int i = i0 + rd * kWarps + warp;
    const bool active = i < i1;
    if (active)
      for (int d = lane; d < p.hd; d += 32) myQ[d] = psg_ld(Q + (long long)i * p.ldq + d) * p.scale;
    float m = -INFINITY, l = 0.f;
    float acc[MAXW][EPW];
#pragma unroll
    for (int w = 0; w < MAXW; ++w)
#pragma unroll
      for (int e = 0; e < EPW; ++e) acc[w][e] = 0.f;

    const int Lk_eff = p.key_len ? max(1, min(p.Lk, p.key_len[b])) : p.Lk;      // keys past a sample's length are masked out
    for (int c0 = 0; c0 < Lk_eff; c0 += p.kc) {
      const int nk = min(p.kc, Lk_eff - c0);
      __syncthreads();  // previous chunk fully consumed
      load_rows<T>(sK, K, p.ldk, c0, nk, p.hd, p.pitch);
      load_rows<T>(sV, V, p.ldv, c0, nk, p.hd, p.pitch);
      __syncthreads();
      if (!active) continue;
      // scores for this chunk
      float cmax = -INFINITY;
      for (int j = lane; j < nk; j += 32) {
        const float s = dot_row<T>(myQ, sK + j * p.pitch, words);
        myP[j] = s;
        cmax = fmaxf(cmax, s);
      }
      cmax = warp_max(cmax);
      const float mnew = fmaxf(m, cmax);
      const float corr = (m == -INFINITY) ? 0.f : __expf(m - mnew);
      float csum = 0.f;
      for (int j = lane; j < nk; j += 32) {
        const float pj = __expf(myP[j] - mnew);
        csum += pj;
        myP[j] = keep(p, b, h, i, c0 + j) ? pj * p.drop_scale : 0.f;
      }
      csum = psg_warp_sum(csum);
      l = l * corr + csum;
      m = mnew;
      __syncwarp();
#pragma unroll
      for (int w = 0; w < MAXW; ++w)
#pragma unroll
        for (int e = 0; e < EPW; ++e) acc[w][e] *= corr;
      for (int j = 0; j < nk; ++j) {
        const float pj = myP[j];
        const uint32_t* vr = sV + j * p.pitch;
#pragma unroll
        for (int w = 0; w < MAXW; ++w) {
          const int wi = lane + 32 * w;
          if (wi < words) {
            float a, bb;
            WordTraits<T>::unpack(vr[wi], a, bb);
            acc[w][0] = fmaf(pj, a, acc[w][0]);
            if (EPW == 2) acc[w][EPW - 1] = fmaf(pj, bb, acc[w][EPW - 1]);
          }
        }
      }
      __syncwarp();
    }
    if (active) {
      const float inv = 1.f / l;
#pragma unroll
      for (int w = 0; w < MAXW; ++w) {
        const int wi = lane + 32 * w;
        if (wi < words) {
#pragma unroll
          for (int e = 0; e < EPW; ++e) psg_st(O + (long long)i * p.ldo + wi * EPW + e, acc[w][e] * inv);
        }
      }
      if (lane == 0 && p.lse) p.lse[((long long)b * p.H + h) * p.Lq + i] = m + __logf(l);
    }
  }
}

// ---------------------------------------------------------------------------------------------------------
// backward pass 1: dQ (and D_i).  Same structure as forward; K/V chunks in smem.
// ---------------------------------------------------------------------------------------------------------
template <typename T>
__global__ void __launch_bounds__(kThreads) attn_bwd_dq_kernel(Params p) {
  constexpr int EPW = WordTraits<T>::EPW;
  constexpr int MAXW = kMaxHeadDim / EPW / 32;
  extern __shared__ uint32_t smem[];
  const int words = p.hd / EPW;
  uint32_t* sK = smem;
  uint32_t* sV = sK + p.kc * p.pitch;
  float* sQ = reinterpret_cast<float*>(sV + p.kc * p.pitch);  // [kWarps][hd]  (scaled q)
  float* sDO = sQ + kWarps * p.hd;                             // [kWarps][hd]
  float* sP = sDO + kWarps * p.hd;                             // [kWarps][kc]  dS
  const int bh = blockIdx.x, b = bh / p.H, h = bh - b * p.H;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int rows_per_cta = (p.Lq + gridDim.y - 1) / gridDim.y;
  const int i0 = blockIdx.y * rows_per_cta, i1 = min(p.Lq, i0 + rows_per_cta);
  const T* Q = reinterpret_cast<const T*>(p.q) + (long long)b * p.Lq * p.ldq + h * p.hd;
  const T* K = reinterpret_cast<const T*>(p.k) + (long long)b * p.Lk * p.ldk + h * p.hd;
  const T* V = reinterpret_cast<const T*>(p.v) + (long long)b * p.Lk * p.ldv + h * p.hd;
  const T* O = reinterpret_cast<const T*>(p.o) + (long long)b * p.Lq * p.ldo + h * p.hd;
  const T* DO = reinterpret_cast<const T*>(p.dout) + (long long)b * p.Lq * p.lddo + h * p.hd;
  T* DQ = reinterpret_cast<T*>(p.dq) + (long long)b * p.Lq * p.lddq + h * p.hd;
  float* myQ = sQ + warp * p.hd;
  float* myDO = sDO + warp * p.hd;
  float* myP = sP + warp * p.kc;
  const int nrounds = (i1 - i0 + kWarps - 1) / kWarps;

  for (int rd = 0; rd < nrounds; ++rd) {
    const int i = i0 + rd * kWarps + warp;
    const bool active = i < i1;
    float lse = 0.f, Di = 0.f;
    if (active) {
      float part = 0.f;
      for (int d = lane; d < p.hd; d += 32) {
        myQ[d] = psg_ld(Q + (long long)i * p.ldq + d) * p.scale;
        const float g = psg_ld(DO + (long long)i * p.lddo + d);
        myDO[d] = g;
        part = fmaf(g, psg_ld(O + (long long)i * p.ldo + d), part);
      }
      Di = psg_warp_sum(part);
      lse = p.lse[((long long)b * p.H + h) * p.Lq + i];
      if (lane == 0) p.dsum[((long long)b * p.H + h) * p.Lq + i] = Di;
    }
    float acc[MAXW][EPW];
#pragma unroll
    for (int w = 0; w < MAXW; ++w)
#pragma unroll
      for (int e = 0; e < EPW; ++e) acc[w][e] = 0.f;
    for (int c0 = 0; c0 < p.Lk; c0 += p.kc) {
      const int nk = min(p.kc, p.Lk - c0);
      __syncthreads();
      load_rows<T>(sK, K, p.ldk, c0, nk, p.hd, p.pitch);
      load_rows<T>(sV, V, p.ldv, c0, nk, p.hd, p.pitch);
      __syncthreads();
      if (!active) continue;
      for (int j = lane; j < nk; j += 32) {
        const float s = dot_row<T>(myQ, sK + j * p.pitch, words);
        const float pr = __expf(s - lse);
        float dp = dot_row<T>(myDO, sV + j * p.pitch, words);
        dp = keep(p, b, h, i, c0 + j) ? dp * p.drop_scale : 0.f;
        myP[j] = pr * (dp - Di);
      }
      __syncwarp();
      for (int j = 0; j < nk; ++j) {
        const float ds = myP[j];
        const uint32_t* kr = sK + j * p.pitch;
#pragma unroll
        for (int w = 0; w < MAXW; ++w) {
          const int wi = lane + 32 * w;
          if (wi < words) {
            float a, bb;
            WordTraits<T>::unpack(kr[wi], a, bb);
            acc[w][0] = fmaf(ds, a, acc[w][0]);
            if (EPW == 2) acc[w][EPW - 1] = fmaf(ds, bb, acc[w][EPW - 1]);
          }
        }
      }
      __syncwarp();
    }
    if (active) {
#pragma unroll
      for (int w = 0; w < MAXW; ++w) {
        const int wi = lane + 32 * w;
        if (wi < words) {
#pragma unroll
          for (int e = 0; e < EPW; ++e) psg_st(DQ + (long long)i * p.lddq + wi * EPW + e, acc[w][e] * p.scale);
        }
      }
    }
  }
}

// ---------------------------------------------------------------------------------------------------------
// backward pass 2: dK, dV.  One warp per key row; Q / dO chunks (over queries) in smem.
// ---------------------------------------------------------------------------------------------------------
template <typename T>
__global__ void __launch_bounds__(kThreads) attn_bwd_dkv_kernel(Params p) {
  constexpr int EPW = WordTraits<T>::EPW;
  constexpr int MAXW = kMaxHeadDim / EPW / 32;
  extern __shared__ uint32_t smem[];
  const int words = p.hd / EPW;
  uint32_t* sQ = smem;                                          // [kc][pitch] raw q rows
  uint32_t* sDO = sQ + p.kc * p.pitch;                          // [kc][pitch]
  float* sKr = reinterpret_cast<float*>(sDO + p.kc * p.pitch);  // [kWarps][hd] (scaled k)
  float* sVr = sKr + kWarps * p.hd;                             // [kWarps][hd]
  float* sP = sVr + kWarps * p.hd;                              // [kWarps][2][kc]   p_drop, ds
  float* sL = sP + kWarps * 2 * p.kc;                           // [kc] lse ; [kc] D
  const int bh = blockIdx.x, b = bh / p.H, h = bh - b * p.H;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int rows_per_cta = (p.Lk + gridDim.y - 1) / gridDim.y;
  const int j0 = blockIdx.y * rows_per_cta, j1 = min(p.Lk, j0 + rows_per_cta);
  const T* Q = reinterpret_cast<const T*>(p.q) + (long long)b * p.Lq * p.ldq + h * p.hd;
  const T* K = reinterpret_cast<const T*>(p.k) + (long long)b * p.Lk * p.ldk + h * p.hd;
  const T* V = reinterpret_cast<const T*>(p.v) + (long long)b * p.Lk * p.ldv + h * p.hd;
  const T* DO = reinterpret_cast<const T*>(p.dout) + (long long)b * p.Lq * p.lddo + h * p.hd;
  T* DK = reinterpret_cast<T*>(p.dk) + (long long)b * p.Lk * p.lddk + h * p.hd;
  T* DV = reinterpret_cast<T*>(p.dv) + (long long)b * p.Lk * p.lddv + h * p.hd;
  float* myK = sKr + warp * p.hd;
  float* myV = sVr + warp * p.hd;
  float* myP = sP + warp * 2 * p.kc;
  float* myDS = myP + p.kc;
  const float* lse_base = p.lse + ((long long)b * p.H + h) * p.Lq;
  const float* d_base = p.dsum + ((long long)b * p.H + h) * p.Lq;
  const int nrounds = (j1 - j0 + kWarps - 1) / kWarps;

  for (int rd = 0; rd < nrounds; ++rd) {
    const int j = j0 + rd * kWarps + warp;
    const bool active = j < j1;
    if (active)
      for (int d = lane; d < p.hd; d += 32) {
        myK[d] = psg_ld(K + (long long)j * p.ldk + d) * p.scale;
        myV[d] = psg_ld(V + (long long)j * p.ldv + d);
      }
    float dk[MAXW][EPW], dv[MAXW][EPW];
#pragma unroll
    for (int w = 0; w < MAXW; ++w)
#pragma unroll
      for (int e = 0; e < EPW; ++e) { dk[w][e] = 0.f; dv[w][e] = 0.f; }
    for (int c0 = 0; c0 < p.Lq; c0 += p.kc) {
      const int nq = min(p.kc, p.Lq - c0);
      __syncthreads();
      load_rows<T>(sQ, Q, p.ldq, c0, nq, p.hd, p.pitch);
      load_rows<T>(sDO, DO, p.lddo, c0, nq, p.hd, p.pitch);
      for (int t = threadIdx.x; t < nq; t += kThreads) { sL[t] = lse_base[c0 + t]; sL[p.kc + t] = d_base[c0 + t]; }
      __syncthreads();
      if (!active) continue;
      for (int i = lane; i < nq; i += 32) {
        const float s = dot_row<T>(myK, sQ + i * p.pitch, words);      // scale folded into myK
        const float pr = __expf(s - sL[i]);
        const bool kp = keep(p, b, h, c0 + i, j);
        float dp = dot_row<T>(myV, sDO + i * p.pitch, words);
        dp = kp ? dp * p.drop_scale : 0.f;
        myP[i] = kp ? pr * p.drop_scale : 0.f;
        myDS[i] = pr * (dp - sL[p.kc + i]);
      }
      __syncwarp();
      for (int i = 0; i < nq; ++i) {
        const float pd = myP[i], ds = myDS[i];
        const uint32_t* qr = sQ + i * p.pitch;
        const uint32_t* gr = sDO + i * p.pitch;
#pragma unroll
        for (int w = 0; w < MAXW; ++w) {
          const int wi = lane + 32 * w;
          if (wi < words) {
            float qa, qb, ga, gb;
            WordTraits<T>::unpack(qr[wi], qa, qb);
            WordTraits<T>::unpack(gr[wi], ga, gb);
            dk[w][0] = fmaf(ds, qa, dk[w][0]);
            dv[w][0] = fmaf(pd, ga, dv[w][0]);
            if (EPW == 2) { dk[w][EPW - 1] = fmaf(ds, qb, dk[w][EPW - 1]); dv[w][EPW - 1] = fmaf(pd, gb, dv[w][EPW - 1]); }
          }
        }
      }
      __syncwarp();
    }
    if (active) {
#pragma unroll
      for (int w = 0; w < MAXW; ++w) {
        const int wi = lane + 32 * w;
        if (wi < words) {
#pragma unroll
          for (int e = 0; e < EPW; ++e) {
            psg_st(DK + (long long)j * p.lddk + wi * EPW + e, dk[w][e] * p.scale);
            psg_st(DV + (long long)j * p.lddv + wi * EPW + e, dv[w][e]);
          }
        }
      }
    }
  }
}

struct Plan { int kc, pitch, ytiles; size_t smem; };

static Plan make_plan(int rows_other, int rows_self, int hd, int epw, int extra_vecs, int bh, size_t per_row_extra_floats) {
  Plan pl;
  int words = hd / epw;
  pl.pitch = (words % 2 == 0) ? words + 1 : words;
  const size_t budget = 160 * 1024;
  size_t fixed = (size_t)kWarps * hd * 4 * extra_vecs;
  int kc = rows_other;
  while (kc > 32) {
    size_t need = (size_t)2 * kc * pl.pitch * 4 + fixed + (size_t)kc * per_row_extra_floats * 4;
    if (need <= budget) break;
    kc = (kc + 1) / 2;
  }
  if (kc < 1) kc = 1;
  pl.kc = kc;
  pl.smem = (size_t)2 * kc * pl.pitch * 4 + fixed + (size_t)kc * per_row_extra_floats * 4;
  int sms = psg_num_sms();
  int yt = (2 * sms + bh - 1) / bh;           // aim for >= 2 CTAs per SM overall
  int max_y = (rows_self + kWarps - 1) / kWarps;
  if (yt > max_y) yt = max_y;
  if (yt < 1) yt = 1;
  pl.ytiles = yt;
  return pl;
}

template <typename K>
static int set_smem(K kern, size_t smem, const char* name) {
  cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(200 * 1024));
  if (e != cudaSuccess) { psg_set_error("%s: cudaFuncSetAttribute: %s", name, cudaGetErrorString(e)); return PSG_ERR_CUDA; }
  (void)smem;
  return PSG_OK;
}

static int check_common(const char* name, int B, int H, int Lq, int Lk, int hd, int dtype) {
  if (B <= 0 || H <= 0 || Lq <= 0 || Lk <= 0) { psg_set_error("%s: bad sizes B=%d H=%d Lq=%d Lk=%d", name, B, H, Lq, Lk); return PSG_ERR_INVALID; }
  if (hd <= 0 || hd > kMaxHeadDim || hd % 2 != 0) { psg_set_error("%s: head_dim %d unsupported (even, <= %d)", name, hd, kMaxHeadDim); return PSG_ERR_UNSUPPORTED; }
  if (dtype != PSG_DTYPE_F32 && dtype != PSG_DTYPE_BF16) { psg_set_error("%s: bad dtype", name); return PSG_ERR_INVALID; }
  if ((long long)B * H > 2147483647LL) { psg_set_error("%s: grid too large", name); return PSG_ERR_INVALID; }
  return PSG_OK;
}

}  // namespace attn

extern "C" {

int psg_attn_fwd_keylen(const void* q, long long ldq, const void* k, long long ldk, const void* v, long long ldv, void* o,
                        long long ldo, float* lse, const int* key_len, int B, int H, int Lq, int Lk, int hd, float scale, int dtype,
                        unsigned long long drop_seed, float drop_p, void* stream);

// o[b, i, h*hd:(h+1)*hd] = softmax_j(q_i . k_j * scale) v_j ; lse saved for backward (may be null for inference).
int psg_attn_fwd(const void* q, long long ldq, const void* k, long long ldk, const void* v, long long ldv, void* o,
                 long long ldo, float* lse, int B, int H, int Lq, int Lk, int hd, float scale, int dtype,
                 unsigned long long drop_seed, float drop_p, void* stream) {
  return psg_attn_fwd_keylen(q, ldq, k, ldk, v, ldv, o, ldo, lse, nullptr, B, H, Lq, Lk, hd, scale, dtype, drop_seed, drop_p, stream);
}

// The same with a per-sample number of valid keys key_len[b] <= Lk (device array, nullable): keys j >= key_len[b] get zero weight --
// BERT's padding mask (src/models/text_encoder.py:152-156: tokenizer padding=True -> attention_mask, a suffix of zeros).
int psg_attn_fwd_keylen(const void* q, long long ldq, const void* k, long long ldk, const void* v, long long ldv, void* o,
                        long long ldo, float* lse, const int* key_len, int B, int H, int Lq, int Lk, int hd, float scale, int dtype,
                        unsigned long long drop_seed, float drop_p, void* stream) {
  using namespace attn;
  int rc = check_common("psg_attn_fwd", B, H, Lq, Lk, hd, dtype);
  if (rc) return rc;
  PSG_CHECK_ARG(q && k && v && o, "psg_attn_fwd: null pointer");
  Params p = {};
  p.key_len = key_len;
  p.q = q; p.k = k; p.v = v; p.o = o; p.lse = lse;
  p.ldq = ldq; p.ldk = ldk; p.ldv = ldv; p.ldo = ldo;
  p.B = B; p.H = H; p.Lq = Lq; p.Lk = Lk; p.hd = hd; p.scale = scale;
  p.drop_seed = drop_seed;
  p.drop_threshold = drop_p > 0.f ? (unsigned int)fmin((double)drop_p * 4294967296.0, 4294967295.0) : 0u;
  p.drop_scale = drop_p > 0.f ? 1.f / (1.f - drop_p) : 1.f;
  const int epw = dtype == PSG_DTYPE_BF16 ? 2 : 1;
  Plan pl = make_plan(Lk, Lq, hd, epw, 1, B * H, 0);
  pl.smem += (size_t)kWarps * pl.kc * 4;
  p.kc = pl.kc; p.pitch = pl.pitch;
  dim3 grid(B * H, pl.ytiles);
  cudaStream_t st = (cudaStream_t)stream;
  if (dtype == PSG_DTYPE_BF16) {
    rc = set_smem(attn_fwd_kernel<__nv_bfloat16>, pl.smem, "psg_attn_fwd"); if (rc) return rc;
    attn_fwd_kernel<__nv_bfloat16><<<grid, kThreads, pl.smem, st>>>(p);
  } else {
    rc = set_smem(attn_fwd_kernel<float>, pl.smem, "psg_attn_fwd"); if (rc) return rc;
    attn_fwd_kernel<float><<<grid, kThreads, pl.smem, st>>>(p);
  }
  PSG_CHECK_LAUNCH("psg_attn_fwd");
  return PSG_OK;
}

// dsum: workspace [B, H, Lq] fp32.
int psg_attn_bwd(const void* q, long long ldq, const void* k, long long ldk, const void* v, long long ldv, const void* o,
                 long long ldo, const void* dout, long long lddo, const float* lse, float* dsum, void* dq, long long lddq,
                 void* dk, long long lddk, void* dv, long long lddv, int B, int H, int Lq, int Lk, int hd, float scale,
                 int dtype, unsigned long long drop_seed, float drop_p, void* stream) {
  using namespace attn;
  int rc = check_common("psg_attn_bwd", B, H, Lq, Lk, hd, dtype);
  if (rc) return rc;
  PSG_CHECK_ARG(q && k && v && o && dout && lse && dsum && dq && dk && dv, "psg_attn_bwd: null pointer");
  Params p = {};
  p.q = q; p.k = k; p.v = v; p.o = const_cast<void*>(o); p.lse = const_cast<float*>(lse);
  p.ldq = ldq; p.ldk = ldk; p.ldv = ldv; p.ldo = ldo;
  p.B = B; p.H = H; p.Lq = Lq; p.Lk = Lk; p.hd = hd; p.scale = scale;
  p.drop_seed = drop_seed;
  p.drop_threshold = drop_p > 0.f ? (unsigned int)fmin((double)drop_p * 4294967296.0, 4294967295.0) : 0u;
  p.drop_scale = drop_p > 0.f ? 1.f / (1.f - drop_p) : 1.f;
  p.dout = dout; p.lddo = lddo; p.dsum = dsum;
  p.dq = dq; p.dk = dk; p.dv = dv; p.lddq = lddq; p.lddk = lddk; p.lddv = lddv;
  const int epw = dtype == PSG_DTYPE_BF16 ? 2 : 1;
  cudaStream_t st = (cudaStream_t)stream;
  {
    Plan pl = make_plan(Lk, Lq, hd, epw, 2, B * H, 0);
    pl.smem += (size_t)kWarps * pl.kc * 4;
    p.kc = pl.kc; p.pitch = pl.pitch;
    dim3 grid(B * H, pl.ytiles);
    if (dtype == PSG_DTYPE_BF16) {
      rc = set_smem(attn_bwd_dq_kernel<__nv_bfloat16>, pl.smem, "psg_attn_bwd"); if (rc) return rc;
      attn_bwd_dq_kernel<__nv_bfloat16><<<grid, kThreads, pl.smem, st>>>(p);
    } else {
      rc = set_smem(attn_bwd_dq_kernel<float>, pl.smem, "psg_attn_bwd"); if (rc) return rc;
      attn_bwd_dq_kernel<float><<<grid, kThreads, pl.smem, st>>>(p);
    }
    PSG_CHECK_LAUNCH("psg_attn_bwd(dq)");
  }
  {
    Plan pl = make_plan(Lq, Lk, hd, epw, 2, B * H, 2);
    pl.smem += (size_t)kWarps * 2 * pl.kc * 4;
    p.kc = pl.kc; p.pitch = pl.pitch;
    dim3 grid(B * H, pl.ytiles);
    if (dtype == PSG_DTYPE_BF16) {
      rc = set_smem(attn_bwd_dkv_kernel<__nv_bfloat16>, pl.smem, "psg_attn_bwd"); if (rc) return rc;
      attn_bwd_dkv_kernel<__nv_bfloat16><<<grid, kThreads, pl.smem, st>>>(p);
    } else {
      rc = set_smem(attn_bwd_dkv_kernel<float>, pl.smem, "psg_attn_bwd"); if (rc) return rc;
      attn_bwd_dkv_kernel<float><<<grid, kThreads, pl.smem, st>>>(p);
    }
    PSG_CHECK_LAUNCH("psg_attn_bwd(dkv)");
  }
  return PSG_OK;
}

}  // extern "C"
