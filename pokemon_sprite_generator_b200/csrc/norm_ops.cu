// GroupNorm (+ optional SiLU) forward / backward over token-major (NHWC) activations.  SURVEY.md §2.1 K4/K5.
//
// Reference ops replaced: nn.GroupNorm + F.silu in ResBlock (src/models/unet.py:79,89,115,127), the eps=1e-6
// norms of CrossAttentionBlock (:156-157,214,231) and final_conv.0 (:397-398), forward and autograd backward.
//
// Layout: x[b, pix, c] with pixel pitch ld (so channel slices of concat buffers work), c contiguous.
// Each thread owns 8 consecutive channels (one 16-byte vector for bf16) for all pixels of its slice, so
// global accesses are coalesced 16B vectors.  Statistics are fp32, reduced in a fixed order:
//   forward : partial (sum, sumsq) per (b, slice, group)          -> apply (normalise, affine, SiLU)
//   backward: partial (sum dn, sum dn*xhat) per (b, slice, chan)  -> apply (dx) + reduce (dgamma, dbeta)
// Bandwidth-bound: forward reads x twice (second pass from L2 when the sample fits) and writes y once.
#include "psg_common.cuh"

namespace {

constexpr int kMaxThreads = 512;

struct GNShape {
  int B, HW, C, G, cpg;
  int vpp;          // vectors (8 ch) per pixel
  int rows;         // pixel rows handled concurrently by a block = blockDim / vpp
  int S;            // slices per sample
  int pix_per_slice;
};

__device__ __forceinline__ float act_fwd(float n, int act) { return act == 1 ? psg_silu(n) : n; }
__device__ __forceinline__ float act_bwd(float n, int act) { return act == 1 ? psg_silu_grad(n) : 1.f; }

template <typename T>
__global__ void gn_stats_kernel(const T* __restrict__ x, long long ld, GNShape s, float* __restrict__ partial) {
  extern __shared__ float sm[];  // [rows*vpp][8] : 4 pair sums + 4 pair sums of squares per thread
  const int b = blockIdx.y, sl = blockIdx.x;
  const int v = threadIdx.x % s.vpp, row = threadIdx.x / s.vpp;
  const int p0 = sl * s.pix_per_slice, p1 = min(s.HW, p0 + s.pix_per_slice);
  if (s.cpg & 1) {
    // odd channels per group (the VAE decoder's GroupNorm(32, 32): one channel per group): per-channel sums,
    // [rows*vpp][16] floats of shared memory
    if (row < s.rows) {
      float sum[8], sq[8];
#pragma unroll
      for (int j = 0; j < 8; ++j) { sum[j] = 0.f; sq[j] = 0.f; }
      const T* base = x + ((long long)b * s.HW) * ld + v * 8;
      for (int p = p0 + row; p < p1; p += s.rows) {
        Vec8<T> t;
        t.load(base + (long long)p * ld);
#pragma unroll
        for (int j = 0; j < 8; ++j) { sum[j] += t.v[j]; sq[j] += t.v[j] * t.v[j]; }
      }
      float* mine = sm + (size_t)threadIdx.x * 16;
#pragma unroll
      for (int j = 0; j < 8; ++j) { mine[j] = sum[j]; mine[8 + j] = sq[j]; }
    }
    __syncthreads();
    float* out = partial + ((long long)(b * s.S + sl) * s.G) * 2;
    for (int g = threadIdx.x; g < s.G; g += blockDim.x) {
      float su = 0.f, sq = 0.f;
      for (int r = 0; r < s.rows; ++r)
        for (int c = g * s.cpg; c < (g + 1) * s.cpg; ++c) {
          const float* src = sm + (size_t)(r * s.vpp + (c >> 3)) * 16;
          su += src[c & 7];
          sq += src[8 + (c & 7)];
        }
      out[2 * g] = su;
      out[2 * g + 1] = sq;
    }
    return;
  }
  if (row < s.rows) {
    float sum[4] = {0.f, 0.f, 0.f, 0.f}, sq[4] = {0.f, 0.f, 0.f, 0.f};
    const T* base = x + ((long long)b * s.HW) * ld + v * 8;
    for (int p = p0 + row; p < p1; p += s.rows) {
      Vec8<T> t;
      t.load(base + (long long)p * ld);
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        sum[j] += t.v[2 * j] + t.v[2 * j + 1];
        sq[j] += t.v[2 * j] * t.v[2 * j] + t.v[2 * j + 1] * t.v[2 * j + 1];
      }
    }
    float* mine = sm + (size_t)threadIdx.x * 8;
#pragma unroll
    for (int j = 0; j < 4; ++j) { mine[j] = sum[j]; mine[4 + j] = sq[j]; }
  }
  __syncthreads();
  // one thread per group folds its channel pairs over all rows in a fixed order (no atomics -> deterministic)
  float* out = partial + ((long long)(b * s.S + sl) * s.G) * 2;
  for (int g = threadIdx.x; g < s.G; g += blockDim.x) {
    float su = 0.f, sq = 0.f;
    const int pair0 = g * s.cpg / 2, pair1 = (g + 1) * s.cpg / 2;
    for (int r = 0; r < s.rows; ++r)
      for (int pi = pair0; pi < pair1; ++pi) {
        const float* src = sm + (size_t)(r * s.vpp + (pi >> 2)) * 8;
        su += src[pi & 3];
        sq += src[4 + (pi & 3)];
      }
    out[2 * g] = su;
    out[2 * g + 1] = sq;
  }
}

template <typename T>
__global__ void gn_apply_kernel(const T* __restrict__ x, long long ld, T* __restrict__ y, long long ldy,
                                const float* __restrict__ gamma, const float* __restrict__ beta, GNShape s,
                                const float* __restrict__ partial, float* __restrict__ stats, float eps, int act) {
  extern __shared__ float sm[];  // [G][2] mean, rstd
  const int b = blockIdx.y, sl = blockIdx.x;
  for (int g = threadIdx.x; g < s.G; g += blockDim.x) {
    float su = 0.f, sq = 0.f;
    for (int k = 0; k < s.S; ++k) {
      const float* pp = partial + ((long long)(b * s.S + k) * s.G + g) * 2;
      su += pp[0];
      sq += pp[1];
    }
    const float m = (float)s.cpg * (float)s.HW;
    const float mean = su / m;
    const float var = fmaxf(sq / m - mean * mean, 0.f);
    const float rstd = rsqrtf(var + eps);
    sm[2 * g] = mean;
    sm[2 * g + 1] = rstd;
    if (sl == 0) { stats[((long long)b * s.G + g) * 2] = mean; stats[((long long)b * s.G + g) * 2 + 1] = rstd; }
  }
  __syncthreads();
  const int v = threadIdx.x % s.vpp, row = threadIdx.x / s.vpp;
  if (row >= s.rows) return;
  float ga[8], be[8], mu[8], rs[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    const int c = v * 8 + j, g = c / s.cpg;
    ga[j] = gamma[c]; be[j] = beta[c]; mu[j] = sm[2 * g]; rs[j] = sm[2 * g + 1];
  }
  const int p0 = sl * s.pix_per_slice, p1 = min(s.HW, p0 + s.pix_per_slice);
  const T* xb = x + ((long long)b * s.HW) * ld + v * 8;
  T* yb = y + ((long long)b * s.HW) * ldy + v * 8;
  for (int p = p0 + row; p < p1; p += s.rows) {
    Vec8<T> t;
    t.load(xb + (long long)p * ld);
#pragma unroll
    for (int j = 0; j < 8; ++j) t.v[j] = act_fwd((t.v[j] - mu[j]) * rs[j] * ga[j] + be[j], act);
    t.store(yb + (long long)p * ldy);
  }
}

// backward pass 1: per-channel partial sums of dn and dn*xhat, dn = dy * act'(n)
template <typename T>
__global__ void gn_bwd_partial_kernel(const T* __restrict__ dy, long long lddy, const T* __restrict__ x, long long ld,
                                      const float* __restrict__ gamma, const float* __restrict__ beta,
                                      const float* __restrict__ stats, GNShape s, float* __restrict__ partial, int act) {
  extern __shared__ float sm[];  // [rows*vpp][16] : 8 x (sum dn), 8 x (sum dn*xhat) per thread
  const int b = blockIdx.y, sl = blockIdx.x;
  const int v = threadIdx.x % s.vpp, row = threadIdx.x / s.vpp;
  if (row < s.rows) {
    float ga[8], be[8], mu[8], rs[8], a1[8], a2[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const int c = v * 8 + j, g = c / s.cpg;
      ga[j] = gamma[c]; be[j] = beta[c];
      mu[j] = stats[((long long)b * s.G + g) * 2]; rs[j] = stats[((long long)b * s.G + g) * 2 + 1];
      a1[j] = 0.f; a2[j] = 0.f;
    }
    const int p0 = sl * s.pix_per_slice, p1 = min(s.HW, p0 + s.pix_per_slice);
    const T* xb = x + ((long long)b * s.HW) * ld + v * 8;
    const T* db = dy + ((long long)b * s.HW) * lddy + v * 8;
    for (int p = p0 + row; p < p1; p += s.rows) {
      Vec8<T> tx, td;
      tx.load(xb + (long long)p * ld);
      td.load(db + (long long)p * lddy);
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        const float xh = (tx.v[j] - mu[j]) * rs[j];
        const float dn = td.v[j] * act_bwd(xh * ga[j] + be[j], act);
        a1[j] += dn;
        a2[j] += dn * xh;
      }
    }
    float* mine = sm + (size_t)threadIdx.x * 16;
#pragma unroll
    for (int j = 0; j < 8; ++j) { mine[j] = a1[j]; mine[8 + j] = a2[j]; }
  }
  __syncthreads();
  float* out = partial + ((long long)(b * s.S + sl) * s.C) * 2;
  for (int c = threadIdx.x; c < s.C; c += blockDim.x) {
    float s1 = 0.f, s2 = 0.f;
    for (int r = 0; r < s.rows; ++r) {
      const float* src = sm + (size_t)(r * s.vpp + (c >> 3)) * 16;
      s1 += src[c & 7];
      s2 += src[8 + (c & 7)];
    }
    out[2 * c] = s1;
    out[2 * c + 1] = s2;
  }
}

// backward pass 2: dx = rstd * (dn*gamma - (A + xhat*Bs)/m), A = sum_g dn*gamma, Bs = sum_g dn*gamma*xhat
template <typename T>
__global__ void gn_bwd_apply_kernel(const T* __restrict__ dy, long long lddy, const T* __restrict__ x, long long ld,
                                    T* __restrict__ dx, long long lddx, const float* __restrict__ gamma,
                                    const float* __restrict__ beta, const float* __restrict__ stats, GNShape s,
                                    const float* __restrict__ partial, int act, int accumulate) {
  extern __shared__ float sm[];  // [G][2] : A, Bs
  const int b = blockIdx.y, sl = blockIdx.x;
  for (int g = threadIdx.x; g < s.G; g += blockDim.x) {
    float A = 0.f, Bv = 0.f;
    for (int c = g * s.cpg; c < (g + 1) * s.cpg; ++c) {
      float s1 = 0.f, s2 = 0.f;
      for (int k = 0; k < s.S; ++k) {
        const float* pp = partial + ((long long)(b * s.S + k) * s.C + c) * 2;
        s1 += pp[0];
        s2 += pp[1];
      }
      A += gamma[c] * s1;
      Bv += gamma[c] * s2;
    }
    sm[2 * g] = A;
    sm[2 * g + 1] = Bv;
  }
  __syncthreads();
  const int v = threadIdx.x % s.vpp, row = threadIdx.x / s.vpp;
  if (row >= s.rows) return;
  const float inv_m = 1.f / ((float)s.cpg * (float)s.HW);
  float ga[8], be[8], mu[8], rs[8], A[8], Bs[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    const int c = v * 8 + j, g = c / s.cpg;
    ga[j] = gamma[c]; be[j] = beta[c];
    mu[j] = stats[((long long)b * s.G + g) * 2]; rs[j] = stats[((long long)b * s.G + g) * 2 + 1];
    A[j] = sm[2 * g] * inv_m; Bs[j] = sm[2 * g + 1] * inv_m;
  }
  const int p0 = sl * s.pix_per_slice, p1 = min(s.HW, p0 + s.pix_per_slice);
  const T* xb = x + ((long long)b * s.HW) * ld + v * 8;
  const T* db = dy + ((long long)b * s.HW) * lddy + v * 8;
  T* ob = dx + ((long long)b * s.HW) * lddx + v * 8;
  for (int p = p0 + row; p < p1; p += s.rows) {
    Vec8<T> tx, td, to;
    tx.load(xb + (long long)p * ld);
    td.load(db + (long long)p * lddy);
    if (accumulate) to.load(ob + (long long)p * lddx);
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const float xh = (tx.v[j] - mu[j]) * rs[j];
      const float dn = td.v[j] * act_bwd(xh * ga[j] + be[j], act);
      const float d = rs[j] * (dn * ga[j] - A[j] - xh * Bs[j]);
      to.v[j] = accumulate ? to.v[j] + d : d;
    }
    to.store(ob + (long long)p * lddx);
  }
}

// dgamma[c] (=|+=) sum over rows of partial[row][c][1]; dbeta[c] (=|+=) ... [0]   (rows = B*S).
// 32 channels x 8 row-lanes per block; each lane walks rows ty, ty+8, ... and an 8-way fixed-order fold finishes
// (the previous one-thread-per-channel serial walk over B*S rows was latency-bound: 0.2 ms per call).
__global__ void __launch_bounds__(256) gn_param_grad_kernel(const float* __restrict__ partial, int rows, int C,
                                                            float* __restrict__ dgamma, float* __restrict__ dbeta, int accumulate) {
  __shared__ float sh[8][32][2];
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
  const int c = blockIdx.x * 32 + tx;
  float s1 = 0.f, s2 = 0.f;
  if (c < C) {
    int r = ty;
    for (; r + 24 < rows; r += 32) {
      const float2 a = *reinterpret_cast<const float2*>(partial + ((long long)r * C + c) * 2);
      const float2 b = *reinterpret_cast<const float2*>(partial + ((long long)(r + 8) * C + c) * 2);
      const float2 d = *reinterpret_cast<const float2*>(partial + ((long long)(r + 16) * C + c) * 2);
      const float2 e = *reinterpret_cast<const float2*>(partial + ((long long)(r + 24) * C + c) * 2);
      s1 += (a.x + b.x) + (d.x + e.x);
      s2 += (a.y + b.y) + (d.y + e.y);
    }
    for (; r < rows; r += 8) {
      const float2 a = *reinterpret_cast<const float2*>(partial + ((long long)r * C + c) * 2);
      s1 += a.x;
      s2 += a.y;
    }
  }
  sh[ty][tx][0] = s1;
  sh[ty][tx][1] = s2;
  __syncthreads();
  if (ty == 0 && c < C) {
    float t1 = 0.f, t2 = 0.f;
#pragma unroll
    for (int k = 0; k < 8; ++k) { t1 += sh[k][tx][0]; t2 += sh[k][tx][1]; }
    dbeta[c] = accumulate ? dbeta[c] + t1 : t1;
    dgamma[c] = accumulate ? dgamma[c] + t2 : t2;
  }
}

int make_shape(GNShape& s, int B, int HW, int C, int G, int slices, int* threads) {
  if (B <= 0 || HW <= 0 || C <= 0 || G <= 0 || C % G != 0 || C % 8 != 0) return -1;
  s.B = B; s.HW = HW; s.C = C; s.G = G; s.cpg = C / G;
  s.vpp = C / 8;
  if (s.vpp > kMaxThreads) return -1;
  s.rows = kMaxThreads / s.vpp;
  if (s.rows > HW) s.rows = HW;
  *threads = ((s.rows * s.vpp + 31) / 32) * 32;
  s.S = slices < 1 ? 1 : slices;
  if (s.S > HW) s.S = HW;
  s.pix_per_slice = (HW + s.S - 1) / s.S;
  s.S = (HW + s.pix_per_slice - 1) / s.pix_per_slice;
  return 0;
}

// ---- LayerNorm over the last dimension, one warp per row (text encoder: src/models/text_encoder.py:158-163 and the
// BertModel layers inside it); statistics in fp32, two passes over a row held in registers (D <= 1024) ----
template <typename TI, typename TO>
__global__ void __launch_bounds__(256) layernorm_kernel(const TI* __restrict__ x, long long ldx, TO* __restrict__ y, long long ldy,
                                                        const float* __restrict__ gamma, const float* __restrict__ beta, long long rows,
                                                        int D, float eps) {
  const long long row = (long long)blockIdx.x * 8 + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  if (row >= rows) return;
  float v[32];                      // D <= 1024: element lane + 32 * i
  float sum = 0.f;
#pragma unroll
  for (int i = 0; i < 32; ++i) {
    const int c = lane + 32 * i;
    v[i] = c < D ? psg_ld(x + row * ldx + c) : 0.f;
    sum += v[i];
  }
  const float mean = psg_warp_sum(sum) / (float)D;
  float sq = 0.f;
#pragma unroll
  for (int i = 0; i < 32; ++i) {
    const int c = lane + 32 * i;
    const float d = c < D ? v[i] - mean : 0.f;
    sq += d * d;
  }
  const float rstd = rsqrtf(psg_warp_sum(sq) / (float)D + eps);
#pragma unroll
  for (int i = 0; i < 32; ++i) {
    const int c = lane + 32 * i;
    if (c < D) psg_st(y + row * ldy + c, (v[i] - mean) * rstd * gamma[c] + beta[c]);
  }
}

// out[row, :] = word[ids[row], :] + pos[row % L, :] + type[type_ids ? type_ids[row] : 0, :]   (BertEmbeddings before its LayerNorm)
__global__ void __launch_bounds__(256) bert_embed_kernel(const long long* __restrict__ ids, const long long* __restrict__ type_ids,
                                                         const float* __restrict__ word, const float* __restrict__ pos,
                                                         const float* __restrict__ type, float* __restrict__ out, long long rows, int L,
                                                         int D, int vocab) {
  const long long row = (long long)blockIdx.x * 8 + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  if (row >= rows) return;
  long long id = ids[row];
  id = id < 0 ? 0 : (id >= vocab ? vocab - 1 : id);
  const long long tt = type_ids ? type_ids[row] : 0;
  const float* w = word + id * D;
  const float* pp = pos + (row % L) * D;
  const float* t = type + tt * D;
  for (int c = lane; c < D; c += 32) out[row * D + c] = w[c] + pp[c] + t[c];
}

}  // namespace

extern "C" {

// y = LayerNorm(x) * gamma + beta over the last dimension (D <= 1024); in / out dtypes independent (PSG_DTYPE_*)
int psg_layernorm(const void* x, long long ldx, void* y, long long ldy, const float* gamma, const float* beta, long long rows, int D,
                  float eps, int in_dtype, int out_dtype, void* stream) {
  if (rows <= 0) return PSG_OK;
  PSG_CHECK_ARG(x && y && gamma && beta && D > 0 && D <= 1024, "psg_layernorm: bad arguments (D=%d, need 1..1024)", D);
  const unsigned grid = (unsigned)((rows + 7) / 8);
  cudaStream_t st = (cudaStream_t)stream;
  if (in_dtype == PSG_DTYPE_F32 && out_dtype == PSG_DTYPE_F32)
    layernorm_kernel<float, float><<<grid, 256, 0, st>>>((const float*)x, ldx, (float*)y, ldy, gamma, beta, rows, D, eps);
  else if (in_dtype == PSG_DTYPE_F32 && out_dtype == PSG_DTYPE_BF16)
    layernorm_kernel<float, __nv_bfloat16><<<grid, 256, 0, st>>>((const float*)x, ldx, (__nv_bfloat16*)y, ldy, gamma, beta, rows, D, eps);
  else if (in_dtype == PSG_DTYPE_BF16 && out_dtype == PSG_DTYPE_BF16)
    layernorm_kernel<__nv_bfloat16, __nv_bfloat16><<<grid, 256, 0, st>>>((const __nv_bfloat16*)x, ldx, (__nv_bfloat16*)y, ldy, gamma, beta, rows, D, eps);
  else if (in_dtype == PSG_DTYPE_BF16 && out_dtype == PSG_DTYPE_F32)
    layernorm_kernel<__nv_bfloat16, float><<<grid, 256, 0, st>>>((const __nv_bfloat16*)x, ldx, (float*)y, ldy, gamma, beta, rows, D, eps);
  else { psg_set_error("psg_layernorm: bad dtype"); return PSG_ERR_INVALID; }
  PSG_CHECK_LAUNCH("psg_layernorm");
  return PSG_OK;
}

// BertEmbeddings sum (before its LayerNorm): ids / type_ids int64 [rows = B * L] (type_ids nullable), tables fp32, out fp32 [rows, D]
int psg_bert_embed(const long long* ids, const long long* type_ids, const float* word, const float* pos, const float* type, float* out,
                   long long rows, int L, int D, int vocab, void* stream) {
  if (rows <= 0) return PSG_OK;
  PSG_CHECK_ARG(ids && word && pos && type && out && L > 0 && D > 0 && vocab > 0, "psg_bert_embed: bad arguments");
  bert_embed_kernel<<<(unsigned)((rows + 7) / 8), 256, 0, (cudaStream_t)stream>>>(ids, type_ids, word, pos, type, out, rows, L, D, vocab);
  PSG_CHECK_LAUNCH("psg_bert_embed");
  return PSG_OK;
}

// Number of slices per sample psg_groupnorm_* will use for (B, HW); workspace sizes derive from it:
//   forward  workspace floats >= B * S * G * 2 ; backward workspace floats >= B * S * C * 2
int psg_groupnorm_slices(int B, int HW) {
  int target = 2 * psg_num_sms();
  int S = (target + B - 1) / (B > 0 ? B : 1);
  int max_s = HW / 16 > 0 ? HW / 16 : 1;   // keep >= 16 pixels per slice
  if (S > max_s) S = max_s;
  if (S < 1) S = 1;
  int pps = (HW + S - 1) / S;
  return (HW + pps - 1) / pps;
}

// y = act(GroupNorm(x)); stats[b][g] = (mean, rstd) saved for backward.  act: 0 none, 1 SiLU.
int psg_groupnorm_fwd(const void* x, long long ld_x, void* y, long long ld_y, const float* gamma, const float* beta,
                      float* stats, float* workspace, int B, int HW, int C, int G, float eps, int act, int dtype,
                      void* stream) {
  PSG_CHECK_ARG(x && y && gamma && beta && stats && workspace, "psg_groupnorm_fwd: null pointer");
  GNShape s;
  int threads;
  PSG_CHECK_ARG(make_shape(s, B, HW, C, G, psg_groupnorm_slices(B, HW), &threads) == 0,
                "psg_groupnorm_fwd: unsupported shape B=%d HW=%d C=%d G=%d (need C%%8==0)", B, HW, C, G);
  PSG_CHECK_ARG(ld_x % 8 == 0 && ld_y % 8 == 0 && ((uintptr_t)x % 16 == 0) && ((uintptr_t)y % 16 == 0),
                "psg_groupnorm_fwd: pitches/pointers must be 16B aligned");
  PSG_CHECK_ARG(B <= 65535, "psg_groupnorm_fwd: B too large");
  dim3 grid(s.S, B);
  cudaStream_t st = (cudaStream_t)stream;
  size_t smem = 2 * G * sizeof(float);
  size_t smem_stats = (size_t)threads * ((s.cpg & 1) ? 16 : 8) * sizeof(float);
  if (dtype == PSG_DTYPE_BF16) {
    gn_stats_kernel<__nv_bfloat16><<<grid, threads, smem_stats, st>>>((const __nv_bfloat16*)x, ld_x, s, workspace);
    gn_apply_kernel<__nv_bfloat16><<<grid, threads, smem, st>>>((const __nv_bfloat16*)x, ld_x, (__nv_bfloat16*)y, ld_y, gamma, beta,
                                                                s, workspace, stats, eps, act);
  } else {
    gn_stats_kernel<float><<<grid, threads, smem_stats, st>>>((const float*)x, ld_x, s, workspace);
    gn_apply_kernel<float><<<grid, threads, smem, st>>>((const float*)x, ld_x, (float*)y, ld_y, gamma, beta, s, workspace, stats,
                                                        eps, act);
  }
  PSG_CHECK_LAUNCH("psg_groupnorm_fwd");
  g_psg_launch_count += 1;  // two kernels
  return PSG_OK;
}

// dx (+)= dL/dx; dgamma/dbeta (+)= parameter grads (fp32).  workspace floats >= B*S*C*2.
int psg_groupnorm_bwd(const void* dy, long long ld_dy, const void* x, long long ld_x, void* dx, long long ld_dx,
                      const float* gamma, const float* beta, const float* stats, float* dgamma, float* dbeta,
                      float* workspace, int B, int HW, int C, int G, int act, int dtype, int accumulate_dx,
                      int accumulate_params, void* stream) {
  PSG_CHECK_ARG(dy && x && dx && gamma && beta && stats && dgamma && dbeta && workspace, "psg_groupnorm_bwd: null pointer");
  GNShape s;
  int threads;
  PSG_CHECK_ARG(make_shape(s, B, HW, C, G, psg_groupnorm_slices(B, HW), &threads) == 0,
                "psg_groupnorm_bwd: unsupported shape B=%d HW=%d C=%d G=%d", B, HW, C, G);
  PSG_CHECK_ARG(s.cpg % 2 == 0, "psg_groupnorm_bwd: odd channels per group are forward-only (C=%d G=%d)", C, G);
  PSG_CHECK_ARG(ld_x % 8 == 0 && ld_dy % 8 == 0 && ld_dx % 8 == 0, "psg_groupnorm_bwd: pitches must be multiples of 8");
  PSG_CHECK_ARG(B <= 65535, "psg_groupnorm_bwd: B too large");
  dim3 grid(s.S, B);
  cudaStream_t st = (cudaStream_t)stream;
  size_t smem1 = (size_t)threads * 16 * sizeof(float), smem2 = 2 * G * sizeof(float);
  if (dtype == PSG_DTYPE_BF16) {
    gn_bwd_partial_kernel<__nv_bfloat16><<<grid, threads, smem1, st>>>((const __nv_bfloat16*)dy, ld_dy, (const __nv_bfloat16*)x, ld_x,
                                                                      gamma, beta, stats, s, workspace, act);
    gn_bwd_apply_kernel<__nv_bfloat16><<<grid, threads, smem2, st>>>((const __nv_bfloat16*)dy, ld_dy, (const __nv_bfloat16*)x, ld_x,
                                                                    (__nv_bfloat16*)dx, ld_dx, gamma, beta, stats, s, workspace, act,
                                                                    accumulate_dx);
  } else {
    gn_bwd_partial_kernel<float><<<grid, threads, smem1, st>>>((const float*)dy, ld_dy, (const float*)x, ld_x, gamma, beta, stats, s,
                                                              workspace, act);
    gn_bwd_apply_kernel<float><<<grid, threads, smem2, st>>>((const float*)dy, ld_dy, (const float*)x, ld_x, (float*)dx, ld_dx, gamma,
                                                            beta, stats, s, workspace, act, accumulate_dx);
  }
  gn_param_grad_kernel<<<(C + 31) / 32, 256, 0, st>>>(workspace, B * s.S, C, dgamma, dbeta, accumulate_params);
  PSG_CHECK_LAUNCH("psg_groupnorm_bwd");
  g_psg_launch_count += 2;  // three kernels
  return PSG_OK;
}

}  // extern "C"
