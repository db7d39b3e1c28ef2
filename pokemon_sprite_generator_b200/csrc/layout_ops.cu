// Small bandwidth-bound kernels around the GEMM engines: layout conversion, strided copies/adds, column sums
// (bias / conditioning gradients), bilinear resize, sinusoidal timestep embedding, token mean-pool, weight
// packing and wgrad finalisation.  All token-major ("NHWC") tensors are addressed as (ptr, pitch) so channel
// slices of concat buffers work everywhere.
//
// Reference ops replaced:
//   nn.Upsample(bilinear, align_corners=False)      src/models/unet.py:365,375,385
//   torch.cat([x, skip], 1)                         src/models/unet.py:482,489,496,503
//   TimestepEmbedding sin/cos                       src/models/unet.py:47-50
//   AdaptiveAvgPool1d(1) over tokens                src/models/unet.py:322,445
//   bias / broadcast-add gradients of conv+linear   autograd of unet.py:116-124
#include "psg_common.cuh"

namespace {

constexpr int kThreads = 256;

inline int blocks_for(long long items, int cap_mult = 8) {
  long long g = (items + kThreads - 1) / kThreads;
  long long cap = (long long)psg_num_sms() * cap_mult;
  if (g > cap) g = cap;
  if (g < 1) g = 1;
  return (int)g;
}

// ---- NCHW fp32 <-> token-major T -------------------------------------------------------------------------
template <typename T>
__global__ void nchw_to_tokens_kernel(const float* __restrict__ src, T* __restrict__ dst, long long ld, int B, int C, int HW) {
  const long long n = (long long)B * C * HW;
  for (long long i = (long long)blockIdx.x * kThreads + threadIdx.x; i < n; i += (long long)gridDim.x * kThreads) {
    const int c = (int)(i % C);
    const long long t = i / C;  // b*HW + pix
    const int pix = (int)(t % HW);
    const int b = (int)(t / HW);
    psg_st(dst + t * ld + c, src[((long long)b * C + c) * HW + pix]);
  }
}
template <typename T>
__global__ void tokens_to_nchw_kernel(const T* __restrict__ src, long long ld, float* __restrict__ dst, int B, int C, int HW) {
  const long long n = (long long)B * C * HW;
  for (long long i = (long long)blockIdx.x * kThreads + threadIdx.x; i < n; i += (long long)gridDim.x * kThreads) {
    const int pix = (int)(i % HW);
    const long long t = i / HW;  // b*C + c
    const int c = (int)(t % C);
    const int b = (int)(t / C);
    dst[i] = psg_ld(src + ((long long)b * HW + pix) * ld + c);
  }
}

// ---- strided copy / add of [rows, C] channel slices (8-wide vectors) -------------------------------------
template <typename T>
__global__ void copy_strided_kernel(const T* __restrict__ src, long long lds, T* __restrict__ dst, long long ldd, long long rows,
                                    int C, int accumulate) {
  const int vpr = C / 8;
  const long long n = rows * vpr;
  for (long long i = (long long)blockIdx.x * kThreads + threadIdx.x; i < n; i += (long long)gridDim.x * kThreads) {
    const long long r = i / vpr;
    const int v = (int)(i - r * vpr);
    Vec8<T> a;
    a.load(src + r * lds + v * 8);
    if (accumulate) {
      Vec8<T> b;
      b.load(dst + r * ldd + v * 8);
#pragma unroll
      for (int j = 0; j < 8; ++j) a.v[j] += b.v[j];
    }
    a.store(dst + r * ldd + v * 8);
  }
}

// ---- column sums: x[groups*rows_per_group, C] -> partial[groups][S][C] ------------------------------------
template <typename T>
__global__ void colsum_partial_kernel(const T* __restrict__ x, long long ld, int rows_per_group, int C, int S,
                                      float* __restrict__ partial) {
  extern __shared__ float sm[];  // [rows*vpp][8] per-thread sums, folded in a fixed order (deterministic)
  const int g = blockIdx.y, sl = blockIdx.x;
  const int vpp = C / 8;
  const int rows = blockDim.x / vpp;
  const int v = threadIdx.x % vpp, row = threadIdx.x / vpp;
  const int per = (rows_per_group + S - 1) / S;
  const int r0 = sl * per, r1 = min(rows_per_group, r0 + per);
  if (row < rows) {
    float acc[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
    const T* base = x + ((long long)g * rows_per_group) * ld + v * 8;
    int r = r0 + row;
    for (; r + 3 * rows < r1; r += 4 * rows) {      // four independent 16-byte loads in flight per thread
      Vec8<T> t0, t1, t2, t3;
      t0.load(base + (long long)r * ld);
      t1.load(base + (long long)(r + rows) * ld);
      t2.load(base + (long long)(r + 2 * rows) * ld);
      t3.load(base + (long long)(r + 3 * rows) * ld);
#pragma unroll
      for (int j = 0; j < 8; ++j) acc[j] += (t0.v[j] + t1.v[j]) + (t2.v[j] + t3.v[j]);
    }
    for (; r < r1; r += rows) {
      Vec8<T> t;
      t.load(base + (long long)r * ld);
#pragma unroll
      for (int j = 0; j < 8; ++j) acc[j] += t.v[j];
    }
    float* mine = sm + (size_t)threadIdx.x * 8;
#pragma unroll
    for (int j = 0; j < 8; ++j) mine[j] = acc[j];
  }
  __syncthreads();
  float* out = partial + ((long long)(g * S + sl)) * C;
  for (int c = threadIdx.x; c < C; c += blockDim.x) {
    float s = 0.f;
    for (int r = 0; r < rows; ++r) s += sm[(size_t)(r * vpp + (c >> 3)) * 8 + (c & 7)];
    out[c] = s;
  }
}
// out_groups[g][c] (=|+=) scale * sum_s partial[g][s][c] ; out_total[c] (=|+=) sum_g of that.
// 32 channels x 8 group-lanes per block, fixed-order fold (deterministic).
__global__ void __launch_bounds__(256) colsum_final_kernel(const float* __restrict__ partial, int groups, int S, int C,
                                                           float* __restrict__ out_groups, long long ld_groups, int acc_groups,
                                                           float* __restrict__ out_total, int acc_total, float scale) {
  __shared__ float sh[8][32];
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
  const int c = blockIdx.x * 32 + tx;
  float tot = 0.f;
  if (c < C) {
    if (out_groups == nullptr) {
      // total only: all groups*S partial rows are equivalent -> spread them over the 8 row-lanes
      const int rows = groups * S;
      int r = ty;
      for (; r + 24 < rows; r += 32)
        tot += (partial[(long long)r * C + c] + partial[(long long)(r + 8) * C + c]) +
               (partial[(long long)(r + 16) * C + c] + partial[(long long)(r + 24) * C + c]);
      for (; r < rows; r += 8) tot += partial[(long long)r * C + c];
      tot *= scale;
    } else {
      for (int g = ty; g < groups; g += 8) {
        float s = 0.f;
        for (int k = 0; k < S; ++k) s += partial[((long long)(g * S + k)) * C + c];
        s *= scale;
        float* o = out_groups + (long long)g * ld_groups + c;
        *o = acc_groups ? *o + s : s;
        tot += s;
      }
    }
  }
  sh[ty][tx] = tot;
  __syncthreads();
  if (ty == 0 && c < C && out_total) {
    float t = 0.f;
#pragma unroll
    for (int k = 0; k < 8; ++k) t += sh[k][tx];
    out_total[c] = acc_total ? out_total[c] + t : t;
  }
}

// ---- conv weights, tap-major bf16 [Cout][kk][Cin] -> dgrad layout [Cin][kk][Cout], every conv of the model in ONE launch ----
// (dgrad then reads a K-major B operand like fprop does: the transposed-in-place MN-major read cost 10-20% per GEMM)
struct TransposeJobs {
  static constexpr int kMax = 64;
  int n;
  int tile_begin[kMax + 1];          // prefix sums of 64x64 tiles over (cin block, cout block, tap)
  long long src[kMax], dst[kMax];    // element offsets into the two bf16 buffers
  int cout[kMax], cin[kMax], kk[kMax];
};

__global__ void __launch_bounds__(256) conv_weight_transpose_kernel(const __nv_bfloat16* __restrict__ src_base,
                                                                     __nv_bfloat16* __restrict__ dst_base, const TransposeJobs jobs) {
  __shared__ __nv_bfloat16 tile[64][72];
  int j = 0;
  while (j + 1 < jobs.n && (int)blockIdx.x >= jobs.tile_begin[j + 1]) ++j;
  const int t = blockIdx.x - jobs.tile_begin[j];
  const int Cout = jobs.cout[j], Cin = jobs.cin[j], kk = jobs.kk[j];
  const int cib = Cin / 64, cob = Cout / 64;
  const int tap = t / (cib * cob), r = t - tap * (cib * cob);
  const int co0 = (r / cib) * 64, ci0 = (r % cib) * 64;
  const __nv_bfloat16* src = src_base + jobs.src[j];
  __nv_bfloat16* dst = dst_base + jobs.dst[j];
  // 64 rows (co) x 8 vectors of 8 ci
  for (int idx = threadIdx.x; idx < 512; idx += 256) {
    const int row = idx >> 3, v = idx & 7;
    const uint4 val = *reinterpret_cast<const uint4*>(src + ((long long)(co0 + row) * kk + tap) * Cin + ci0 + v * 8);
    *reinterpret_cast<uint4*>(&tile[row][v * 8]) = val;
  }
  __syncthreads();
  // 64 rows (ci) x 8 vectors of 8 co
  for (int idx = threadIdx.x; idx < 512; idx += 256) {
    const int row = idx >> 3, v = idx & 7;
    __align__(16) __nv_bfloat16 out[8];
#pragma unroll
    for (int e = 0; e < 8; ++e) out[e] = tile[v * 8 + e][row];
    *reinterpret_cast<uint4*>(dst + ((long long)(ci0 + row) * kk + tap) * Cout + co0 + v * 8) = *reinterpret_cast<const uint4*>(out);
  }
}

// ---- bilinear resize (align_corners=False), token-major ----------------------------------------------------
__device__ __forceinline__ void src_index(int dst, float scale, int in_size, int& i0, int& i1, float& l1) {
  float s = scale * ((float)dst + 0.5f) - 0.5f;
  if (s < 0.f) s = 0.f;
  i0 = (int)s;
  if (i0 > in_size - 1) i0 = in_size - 1;
  i1 = i0 + (i0 < in_size - 1 ? 1 : 0);
  l1 = s - (float)i0;
}
template <typename T>
__global__ void upsample_fwd_kernel(const T* __restrict__ x, long long ldx, T* __restrict__ y, long long ldy, int B, int C, int IH,
                                    int IW, int OH, int OW) {
  const int vpp = C / 8;
  const long long n = (long long)B * OH * OW * vpp;
  const float sh = (float)IH / (float)OH, sw = (float)IW / (float)OW;
  for (long long i = (long long)blockIdx.x * kThreads + threadIdx.x; i < n; i += (long long)gridDim.x * kThreads) {
    const int v = (int)(i % vpp);
    long long t = i / vpp;
    const int ox = (int)(t % OW); t /= OW;
    const int oy = (int)(t % OH);
    const int b = (int)(t / OH);
    int y0, y1, x0, x1; float ly, lx;
    src_index(oy, sh, IH, y0, y1, ly);
    src_index(ox, sw, IW, x0, x1, lx);
    const T* base = x + ((long long)b * IH * IW) * ldx + v * 8;
    Vec8<T> a, bq, c, d, o;
    a.load(base + ((long long)y0 * IW + x0) * ldx);
    bq.load(base + ((long long)y0 * IW + x1) * ldx);
    c.load(base + ((long long)y1 * IW + x0) * ldx);
    d.load(base + ((long long)y1 * IW + x1) * ldx);
    const float hy = 1.f - ly, hx = 1.f - lx;
#pragma unroll
    for (int j = 0; j < 8; ++j) o.v[j] = hy * (hx * a.v[j] + lx * bq.v[j]) + ly * (hx * c.v[j] + lx * d.v[j]);
    o.store(y + (((long long)b * OH + oy) * OW + ox) * ldy + v * 8);
  }
}
// gather form of the backward: dx[iy,ix] = sum over output pixels that read (iy,ix) of weight * dy
template <typename T>
__global__ void upsample_bwd_kernel(const T* __restrict__ dy, long long lddy, T* __restrict__ dx, long long lddx, int B, int C, int IH,
                                    int IW, int OH, int OW, int accumulate) {
  const int vpp = C / 8;
  const long long n = (long long)B * IH * IW * vpp;
  const float sh = (float)IH / (float)OH, sw = (float)IW / (float)OW;
  for (long long i = (long long)blockIdx.x * kThreads + threadIdx.x; i < n; i += (long long)gridDim.x * kThreads) {
    const int v = (int)(i % vpp);
    long long t = i / vpp;
    const int ix = (int)(t % IW); t /= IW;
    const int iy = (int)(t % IH);
    const int b = (int)(t / IH);
    float acc[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
    // candidate output rows/cols: those whose source window [i0, i1] can contain iy / ix
    const int oy_lo = max(0, (int)floorf(((float)iy - 1.f + 0.5f) / sh - 0.5f) - 1);
    const int oy_hi = min(OH - 1, (int)ceilf(((float)iy + 1.f + 0.5f) / sh - 0.5f) + 1);
    const int ox_lo = max(0, (int)floorf(((float)ix - 1.f + 0.5f) / sw - 0.5f) - 1);
    const int ox_hi = min(OW - 1, (int)ceilf(((float)ix + 1.f + 0.5f) / sw - 0.5f) + 1);
    for (int oy = oy_lo; oy <= oy_hi; ++oy) {
      int y0, y1; float ly;
      src_index(oy, sh, IH, y0, y1, ly);
      float wy = 0.f;
      if (y0 == iy) wy += 1.f - ly;
      if (y1 == iy) wy += ly;
      if (wy == 0.f) continue;
      for (int ox = ox_lo; ox <= ox_hi; ++ox) {
        int x0, x1; float lx;
        src_index(ox, sw, IW, x0, x1, lx);
        float wx = 0.f;
        if (x0 == ix) wx += 1.f - lx;
        if (x1 == ix) wx += lx;
        if (wx == 0.f) continue;
        Vec8<T> g;
        g.load(dy + (((long long)b * OH + oy) * OW + ox) * lddy + v * 8);
        const float w = wy * wx;
#pragma unroll
        for (int j = 0; j < 8; ++j) acc[j] = fmaf(w, g.v[j], acc[j]);
      }
    }
    T* o = dx + (((long long)b * IH + iy) * IW + ix) * lddx + v * 8;
    Vec8<T> r;
    if (accumulate) {
      r.load(o);
#pragma unroll
      for (int j = 0; j < 8; ++j) r.v[j] += acc[j];
    } else {
#pragma unroll
      for (int j = 0; j < 8; ++j) r.v[j] = acc[j];
    }
    r.store(o);
  }
}

// zero-insertion: out[b, 2p, 2q, :] = dy[b, p, q, :], zeros elsewhere (stride-2 dgrad as a stride-1 conv)
template <typename T>
__global__ void dilate2_kernel(const T* __restrict__ dy, long long lddy, T* __restrict__ out, long long ldo, int B, int C, int P, int Q,
                               int H, int W) {
  const int vpp = C / 8;
  const long long n = (long long)B * H * W * vpp;
  for (long long i = (long long)blockIdx.x * kThreads + threadIdx.x; i < n; i += (long long)gridDim.x * kThreads) {
    const int v = (int)(i % vpp);
    long long t = i / vpp;
    const int w = (int)(t % W); t /= W;
    const int h = (int)(t % H);
    const int b = (int)(t / H);
    Vec8<T> r;
#pragma unroll
    for (int j = 0; j < 8; ++j) r.v[j] = 0.f;
    if (!(h & 1) && !(w & 1) && (h >> 1) < P && (w >> 1) < Q) r.load(dy + (((long long)b * P + (h >> 1)) * Q + (w >> 1)) * lddy + v * 8);
    r.store(out + (((long long)b * H + h) * W + w) * ldo + v * 8);
  }
}

// ---- stride-2 dgrad by output parity (3x3, stride 2, pad 1; src/models/unet.py:335-349 downsample convs) --------------------------
// dX[2a+pi, 2b+pj] only receives the taps whose parity matches: row taps r = 1 for pi = 0 (reading dY[a]), r = 2 and 0 for pi = 1
// (reading dY[a] and dY[a+1]); likewise for columns.  Each of the four classes is a 2x2 (or 1x1) stride-1 convolution over dY:
// 13 taps instead of the 36 a zero-inserted stride-1 convolution executes.
// Class weights from the dgrad layout wd [Cin][9][Cout] (tap = 3 r + s):  w00 [Cin][Cout];  w01, w10, w11 [Cin][4][Cout] with
// window tap (dr, ds) at index 2 dr + ds (unused window taps of the 1x2 / 2x1 classes are zero).
template <typename T>
__global__ void dgrad_s2_weights_kernel(const T* __restrict__ wd, T* __restrict__ w00, T* __restrict__ w01, T* __restrict__ w10,
                                        T* __restrict__ w11, int Cin, int Cout) {
  const long long n = (long long)Cin * 4 * Cout;
  for (long long i = (long long)blockIdx.x * kThreads + threadIdx.x; i < n; i += (long long)gridDim.x * kThreads) {
    const int co = (int)(i % Cout);
    const int win = (int)((i / Cout) & 3), dr = win >> 1, ds = win & 1;
    const long long ci = i / (4LL * Cout);
    const T* src = wd + ci * 9 * Cout + co;
    const int r1 = 2 - 2 * dr, s1 = 2 - 2 * ds;             // the tap of an odd output row / column that reads dY[a + dr] / dY[b + ds]
    const T zero = T(0.f);
    w11[i] = src[(3 * r1 + s1) * Cout];
    w01[i] = dr == 0 ? src[(3 * 1 + s1) * Cout] : zero;     // even row (r = 1, reads dY[a] only), odd column
    w10[i] = ds == 0 ? src[(3 * r1 + 1) * Cout] : zero;     // odd row, even column
    if (win == 0) w00[ci * Cout + co] = src[4 * Cout];
  }
}

// dX[b, 2a+pi, 2b'+pj, :] (+)= class_{pi pj}[b, a, b', :] for the positions inside H x W; classes are [B, P, Q, C] (pixel pitch C)
template <typename T>
__global__ void interleave2x2_kernel(const T* __restrict__ c00, const T* __restrict__ c01, const T* __restrict__ c10, const T* __restrict__ c11,
                                     T* __restrict__ out, long long ldo, int B, int C, int P, int Q, int H, int W, int accumulate) {
  const int vpp = C / 8;
  const long long n = (long long)B * H * W * vpp;
  for (long long i = (long long)blockIdx.x * kThreads + threadIdx.x; i < n; i += (long long)gridDim.x * kThreads) {
    const int v = (int)(i % vpp);
    long long t = i / vpp;
    const int w = (int)(t % W); t /= W;
    const int h = (int)(t % H);
    const int b = (int)(t / H);
    const T* src = (h & 1) ? ((w & 1) ? c11 : c10) : ((w & 1) ? c01 : c00);
    Vec8<T> r;
    r.load(src + ((((long long)b * P + (h >> 1)) * Q + (w >> 1)) * C) + v * 8);
    T* dst = out + (((long long)b * H + h) * W + w) * ldo + v * 8;
    if (accumulate) {
      Vec8<T> o;
      o.load(dst);
#pragma unroll
      for (int j = 0; j < 8; ++j) r.v[j] += o.v[j];
    }
    r.store(dst);
  }
}

// out = alpha * dropout(x; seed, threshold) -- backward of an epilogue dropout applied before a GEMM operand
template <typename T>
__global__ void dropout_scale_kernel(const T* __restrict__ x, long long ldx, T* __restrict__ out, long long ldo, long long rows, int C,
                                     float alpha, unsigned long long seed, unsigned int threshold, float keep_scale) {
  const int vpr = C / 8;
  const long long n = rows * vpr;
  for (long long i = (long long)blockIdx.x * kThreads + threadIdx.x; i < n; i += (long long)gridDim.x * kThreads) {
    const long long r = i / vpr;
    const int v = (int)(i - r * vpr);
    Vec8<T> a;
    a.load(x + r * ldx + v * 8);
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const bool kp = threshold == 0 || psg_drop_keep(seed, (uint64_t)(r * C + v * 8 + j), threshold);
      a.v[j] = kp ? a.v[j] * keep_scale * alpha : 0.f;
    }
    a.store(out + r * ldo + v * 8);
  }
}

// ---- conditioning inputs -----------------------------------------------------------------------------------
__global__ void timestep_embedding_kernel(const long long* __restrict__ t, const float* __restrict__ coeff, float* __restrict__ out,
                                          int B, int half) {
  const int n = B * half;
  for (int i = blockIdx.x * kThreads + threadIdx.x; i < n; i += gridDim.x * kThreads) {
    const int b = i / half, k = i - b * half;
    const float e = __fmul_rn((float)t[b], coeff[k]);
    out[(long long)b * 2 * half + k] = sinf(e);
    out[(long long)b * 2 * half + half + k] = cosf(e);
  }
}
__global__ void mean_pool_kernel(const float* __restrict__ x, float* __restrict__ out, int B, int L, int D) {
  const int n = B * D;
  for (int i = blockIdx.x * kThreads + threadIdx.x; i < n; i += gridDim.x * kThreads) {
    const int b = i / D, d = i - b * D;
    float s = 0.f;
    for (int l = 0; l < L; ++l) s += x[((long long)b * L + l) * D + d];
    out[i] = s / (float)L;
  }
}

// ---- weight packing ----------------------------------------------------------------------------------------
// OIHW fp32 -> wp[Cout_p][kk][Cin_p] (fprop B operand) and wd[Cin_p][kk][Cout_p] (dgrad B operand), dtype T.  The packed
// channel counts may exceed the real ones (edge layers padded to the 64-wide tensor-core tile); the padding is zeroed
// once by the caller and never written here.
template <typename T>
__global__ void __launch_bounds__(256) pack_conv_kernel(const float* __restrict__ w, T* __restrict__ wp, T* __restrict__ wd, int Cout,
                                                        int Cin, int kk, int Cin_p, int Cout_p) {
  extern __shared__ float tile[];  // [32 co][32 ci][kk] (+1 pad per ci row)
  const int co0 = blockIdx.y * 32, ci0 = blockIdx.x * 32;
  const int row = 32 * kk + 1;    // floats per co
  // coalesced read: for each co, the (ci, tap) block of 32*kk consecutive floats
  for (int idx = threadIdx.x; idx < 32 * 32 * kk; idx += blockDim.x) {
    const int co = idx / (32 * kk), r = idx - co * (32 * kk);
    const int ci = r / kk;
    float v = 0.f;
    if (co0 + co < Cout && ci0 + ci < Cin) v = w[((long long)(co0 + co) * Cin + ci0) * kk + r];
    tile[co * row + r] = v;
  }
  __syncthreads();
  // wp[co][tap][ci]: 32 consecutive ci per (co, tap)
  if (wp) {
    for (int idx = threadIdx.x; idx < 32 * kk * 32; idx += blockDim.x) {
      const int ci = idx & 31, t = idx >> 5;
      const int tap = t % kk, co = t / kk;
      if (co0 + co < Cout && ci0 + ci < Cin)
        psg_st(wp + ((long long)(co0 + co) * kk + tap) * Cin_p + ci0 + ci, tile[co * row + ci * kk + tap]);
    }
  }
  // wd[ci][tap][co]: 32 consecutive co per (ci, tap)
  if (wd) {
    for (int idx = threadIdx.x; idx < 32 * kk * 32; idx += blockDim.x) {
      const int co = idx & 31, t = idx >> 5;
      const int tap = t % kk, ci = t / kk;
      if (co0 + co < Cout && ci0 + ci < Cin)
        psg_st(wd + ((long long)(ci0 + ci) * kk + tap) * Cout_p + co0 + co, tile[co * row + ci * kk + tap]);
    }
  }
}
// [N][K] fp32 -> wk[N][K] (cast) and wt[K][N] (transpose + cast)
template <typename T>
__global__ void pack_linear_kernel(const float* __restrict__ w, T* __restrict__ wk, T* __restrict__ wt, int N, int K) {
  __shared__ float tile[32][33];
  const int n0 = blockIdx.y * 32, k0 = blockIdx.x * 32;
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;  // 32 x 8
  for (int r = ty; r < 32; r += 8) {
    const int n = n0 + r, k = k0 + tx;
    float v = 0.f;
    if (n < N && k < K) {
      v = w[(long long)n * K + k];
      if (wk) psg_st(wk + (long long)n * K + k, v);
    }
    tile[r][tx] = v;
  }
  __syncthreads();
  if (wt) {
    for (int r = ty; r < 32; r += 8) {
      const int k = k0 + r, n = n0 + tx;
      if (n < N && k < K) psg_st(wt + (long long)k * N + n, tile[tx][r]);
    }
  }
}
// grad_oihw[co][ci][tap] (=|+=) sum_s partial[s][co][tap*Cin_p + ci]; one block per (co, 64-wide ci chunk): coalesced
// reads of the packed layout, smem transpose, coalesced writes of the OIHW layout.
__global__ void __launch_bounds__(256) wgrad_finalize_kernel(const float* __restrict__ partial, int S, long long split_stride,
                                                             float* __restrict__ grad, int Cout, int Cin, int kk, int Cin_p,
                                                             int accumulate) {
  extern __shared__ float tile[];  // [kk][64+1]
  const int co = blockIdx.y, ci0 = blockIdx.x * 64;
  for (int idx = threadIdx.x; idx < kk * 64; idx += blockDim.x) {
    const int tap = idx >> 6, ci = idx & 63;
    float s = 0.f;
    if (ci0 + ci < Cin) {
      const long long src = ((long long)co * kk + tap) * Cin_p + ci0 + ci;
      for (int k = 0; k < S; ++k) s += partial[(long long)k * split_stride + src];
    }
    tile[tap * 65 + ci] = s;
  }
  __syncthreads();
  const int nci = min(64, Cin - ci0);
  float* out = grad + ((long long)co * Cin + ci0) * kk;
  for (int idx = threadIdx.x; idx < nci * kk; idx += blockDim.x) {
    const int ci = idx / kk, tap = idx - ci * kk;
    const float v = tile[tap * 65 + ci];
    out[idx] = accumulate ? out[idx] + v : v;
  }
}
// out (=|+=) sum_s partial[s][i]   (linear wgrad split-K reduction, same layout)
__global__ void sum_partials_kernel(const float* __restrict__ partial, int S, long long split_stride, float* __restrict__ out,
                                    long long n, int accumulate) {
  for (long long i = (long long)blockIdx.x * kThreads + threadIdx.x; i < n; i += (long long)gridDim.x * kThreads) {
    float s = 0.f;
    for (int k = 0; k < S; ++k) s += partial[(long long)k * split_stride + i];
    out[i] = accumulate ? out[i] + s : s;
  }
}

}  // namespace

// ---- per-sample transpose in[b][r][c] -> out[b][c][r] (the VAE decoder's raw-reshape K / V, src/models/vae_decoder.py:54-55) ----
template <typename T>
__global__ void batched_transpose_kernel(const T* __restrict__ in, T* __restrict__ out, int R, int Cc) {
  __shared__ T tile[32][33];
  const long long base = (long long)blockIdx.z * R * Cc;
  const int r0 = blockIdx.y * 32, c0 = blockIdx.x * 32;
  for (int i = threadIdx.y; i < 32; i += blockDim.y) {
    const int r = r0 + i, c = c0 + threadIdx.x;
    if (r < R && c < Cc) tile[i][threadIdx.x] = in[base + (long long)r * Cc + c];
  }
  __syncthreads();
  for (int i = threadIdx.y; i < 32; i += blockDim.y) {
    const int c = c0 + i, r = r0 + threadIdx.x;
    if (r < R && c < Cc) out[base + (long long)c * R + r] = tile[threadIdx.x][i];
  }
}

__global__ void tanh_kernel(const float* __restrict__ x, float* __restrict__ y, size_t n) {
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) y[i] = tanhf(x[i]);
}

#define DISPATCH_T(dtype, ...)                                           \
  if ((dtype) == PSG_DTYPE_BF16) { using T = __nv_bfloat16; __VA_ARGS__; } \
  else if ((dtype) == PSG_DTYPE_F32) { using T = float; __VA_ARGS__; }     \
  else { psg_set_error("bad dtype %d", (int)(dtype)); return PSG_ERR_INVALID; }

extern "C" {

// out[b][c][r] = in[b][r][c], b < B (contiguous [B, R, Cc] in, [B, Cc, R] out)
int psg_batched_transpose(const void* in, void* out, int B, int R, int Cc, int dtype, void* stream) {
  PSG_CHECK_ARG(B >= 0 && R > 0 && Cc > 0 && B <= 65535, "psg_batched_transpose: bad sizes");
  if (B == 0) return PSG_OK;
  PSG_CHECK_ARG(in && out && in != out, "psg_batched_transpose: null or aliased pointer");
  dim3 grid((Cc + 31) / 32, (R + 31) / 32, B), block(32, 8);
  DISPATCH_T(dtype, (batched_transpose_kernel<T><<<grid, block, 0, (cudaStream_t)stream>>>((const T*)in, (T*)out, R, Cc)));
  PSG_CHECK_LAUNCH("psg_batched_transpose");
  return PSG_OK;
}

// y = tanh(x), fp32 (the VAE decoder's output activation, src/models/vae_decoder.py:174); in place allowed
int psg_tanh(const float* x, float* y, long long n, void* stream) {
  if (n <= 0) return PSG_OK;
  PSG_CHECK_ARG(x && y, "psg_tanh: null pointer");
  long long blocks = (n + 255) / 256;
  const long long cap = (long long)psg_num_sms() * 16;
  if (blocks > cap) blocks = cap;
  tanh_kernel<<<(unsigned)blocks, 256, 0, (cudaStream_t)stream>>>(x, y, (size_t)n);
  PSG_CHECK_LAUNCH("psg_tanh");
  return PSG_OK;
}

int psg_nchw_to_tokens(const float* src, void* dst, long long ld, int B, int C, int HW, int dtype, void* stream) {
  PSG_CHECK_ARG(B >= 0 && C > 0 && HW > 0, "psg_nchw_to_tokens: bad sizes");
  if (B == 0) return PSG_OK;
  PSG_CHECK_ARG(src && dst, "psg_nchw_to_tokens: null pointer");
  DISPATCH_T(dtype, (nchw_to_tokens_kernel<T><<<blocks_for((long long)B * C * HW), kThreads, 0, (cudaStream_t)stream>>>(src, (T*)dst, ld, B, C, HW)));
  PSG_CHECK_LAUNCH("psg_nchw_to_tokens");
  return PSG_OK;
}

int psg_tokens_to_nchw(const void* src, long long ld, float* dst, int B, int C, int HW, int dtype, void* stream) {
  PSG_CHECK_ARG(B >= 0 && C > 0 && HW > 0, "psg_tokens_to_nchw: bad sizes");
  if (B == 0) return PSG_OK;
  PSG_CHECK_ARG(src && dst, "psg_tokens_to_nchw: null pointer");
  DISPATCH_T(dtype, (tokens_to_nchw_kernel<T><<<blocks_for((long long)B * C * HW), kThreads, 0, (cudaStream_t)stream>>>((const T*)src, ld, dst, B, C, HW)));
  PSG_CHECK_LAUNCH("psg_tokens_to_nchw");
  return PSG_OK;
}

int psg_copy_strided(const void* src, long long lds, void* dst, long long ldd, long long rows, int C, int accumulate, int dtype,
                     void* stream) {
  PSG_CHECK_ARG(rows >= 0 && C > 0 && C % 8 == 0 && lds % 8 == 0 && ldd % 8 == 0, "psg_copy_strided: C and pitches must be multiples of 8");
  if (rows == 0) return PSG_OK;
  PSG_CHECK_ARG(src && dst, "psg_copy_strided: null pointer");
  DISPATCH_T(dtype, (copy_strided_kernel<T><<<blocks_for(rows * (C / 8)), kThreads, 0, (cudaStream_t)stream>>>((const T*)src, lds, (T*)dst, ldd, rows, C, accumulate)));
  PSG_CHECK_LAUNCH("psg_copy_strided");
  return PSG_OK;
}

// slices used by psg_colsum for a given group size; workspace floats >= 32 + groups * S * C, the first 32 zeroed once
int psg_colsum_slices(int groups, int rows_per_group) {
  int S = (2 * psg_num_sms() + groups - 1) / (groups > 0 ? groups : 1);
  int max_s = rows_per_group / 16 > 0 ? rows_per_group / 16 : 1;
  if (S > max_s) S = max_s;
  return S < 1 ? 1 : S;
}

// Per-group and total column sums (times `scale`) of x[groups*rows_per_group, C]: out_groups[g][c], out_total[c]
// (either may be null).
int psg_colsum(const void* x, long long ld, int groups, int rows_per_group, int C, float* out_groups, long long ld_groups,
               int acc_groups, float* out_total, int acc_total, float scale, float* workspace, int dtype, void* stream) {
  PSG_CHECK_ARG(x && workspace, "psg_colsum: null pointer");
  if (C > 4096 && C % 8 == 0) {
    // wide matrices (the [B, sum Cout] conditioning gradient): column chunks of 4096, same workspace (stream-ordered)
    const size_t esz = dtype == PSG_DTYPE_BF16 ? 2 : 4;
    for (int c0 = 0; c0 < C; c0 += 4096) {
      const int cw = C - c0 < 4096 ? C - c0 : 4096;
      int rc = psg_colsum((const char*)x + (size_t)c0 * esz, ld, groups, rows_per_group, cw, out_groups ? out_groups + c0 : nullptr,
                          ld_groups, acc_groups, out_total ? out_total + c0 : nullptr, acc_total, scale, workspace, dtype, stream);
      if (rc) return rc;
    }
    return PSG_OK;
  }
  PSG_CHECK_ARG(groups > 0 && rows_per_group > 0 && C > 0 && C % 8 == 0 && C / 8 <= 512 && ld % 8 == 0, "psg_colsum: bad sizes (C=%d)", C);
  PSG_CHECK_ARG(groups <= 65535, "psg_colsum: too many groups");
  const int S = psg_colsum_slices(groups, rows_per_group);
  const int vpp = C / 8;
  int rows = 512 / vpp;
  if (rows < 1) rows = 1;
  int threads = ((rows * vpp + 31) / 32) * 32;
  dim3 grid(S, groups);
  cudaStream_t st = (cudaStream_t)stream;
  // (a single-launch variant whose last block folds the slices was measured 2.5x slower: the fold is a serial tail)
  float* partial = workspace + 32;
  DISPATCH_T(dtype, (colsum_partial_kernel<T><<<grid, threads, (size_t)threads * 8 * sizeof(float), st>>>((const T*)x, ld, rows_per_group, C, S, partial)));
  colsum_final_kernel<<<(C + 31) / 32, 256, 0, st>>>(partial, groups, S, C, out_groups, ld_groups, acc_groups, out_total, acc_total, scale);
  PSG_CHECK_LAUNCH("psg_colsum");
  g_psg_launch_count += 1;  // two kernels
  return PSG_OK;
}

int psg_upsample_bilinear_fwd(const void* x, long long ldx, void* y, long long ldy, int B, int C, int IH, int IW, int OH, int OW,
                              int dtype, void* stream) {
  PSG_CHECK_ARG(x && y && B > 0 && C % 8 == 0 && ldx % 8 == 0 && ldy % 8 == 0, "psg_upsample_bilinear_fwd: bad args");
  DISPATCH_T(dtype, (upsample_fwd_kernel<T><<<blocks_for((long long)B * OH * OW * (C / 8)), kThreads, 0, (cudaStream_t)stream>>>((const T*)x, ldx, (T*)y, ldy, B, C, IH, IW, OH, OW)));
  PSG_CHECK_LAUNCH("psg_upsample_bilinear_fwd");
  return PSG_OK;
}

int psg_upsample_bilinear_bwd(const void* dy, long long lddy, void* dx, long long lddx, int B, int C, int IH, int IW, int OH, int OW,
                              int accumulate, int dtype, void* stream) {
  PSG_CHECK_ARG(dy && dx && B > 0 && C % 8 == 0 && lddy % 8 == 0 && lddx % 8 == 0, "psg_upsample_bilinear_bwd: bad args");
  DISPATCH_T(dtype, (upsample_bwd_kernel<T><<<blocks_for((long long)B * IH * IW * (C / 8)), kThreads, 0, (cudaStream_t)stream>>>((const T*)dy, lddy, (T*)dx, lddx, B, C, IH, IW, OH, OW, accumulate)));
  PSG_CHECK_LAUNCH("psg_upsample_bilinear_bwd");
  return PSG_OK;
}

int psg_dilate2(const void* dy, long long lddy, void* out, long long ldo, int B, int C, int P, int Q, int H, int W, int dtype,
                void* stream) {
  PSG_CHECK_ARG(dy && out && B > 0 && C % 8 == 0 && lddy % 8 == 0 && ldo % 8 == 0, "psg_dilate2: bad args");
  DISPATCH_T(dtype, (dilate2_kernel<T><<<blocks_for((long long)B * H * W * (C / 8)), kThreads, 0, (cudaStream_t)stream>>>((const T*)dy, lddy, (T*)out, ldo, B, C, P, Q, H, W)));
  PSG_CHECK_LAUNCH("psg_dilate2");
  return PSG_OK;
}

int psg_dgrad_s2_weights(const void* wd, void* w00, void* w01, void* w10, void* w11, int Cin, int Cout, int dtype, void* stream) {
  PSG_CHECK_ARG(wd && w00 && w01 && w10 && w11 && Cin > 0 && Cout > 0, "psg_dgrad_s2_weights: bad args");
  DISPATCH_T(dtype, (dgrad_s2_weights_kernel<T><<<blocks_for((long long)Cin * 4 * Cout), kThreads, 0, (cudaStream_t)stream>>>(
                        (const T*)wd, (T*)w00, (T*)w01, (T*)w10, (T*)w11, Cin, Cout)));
  PSG_CHECK_LAUNCH("psg_dgrad_s2_weights");
  return PSG_OK;
}

int psg_interleave2x2(const void* c00, const void* c01, const void* c10, const void* c11, void* out, long long ldo, int B, int C, int P,
                      int Q, int H, int W, int accumulate, int dtype, void* stream) {
  PSG_CHECK_ARG(c00 && c01 && c10 && c11 && out && B > 0 && C % 8 == 0 && ldo % 8 == 0 && 2 * P >= H && 2 * Q >= W, "psg_interleave2x2: bad args");
  DISPATCH_T(dtype, (interleave2x2_kernel<T><<<blocks_for((long long)B * H * W * (C / 8)), kThreads, 0, (cudaStream_t)stream>>>(
                        (const T*)c00, (const T*)c01, (const T*)c10, (const T*)c11, (T*)out, ldo, B, C, P, Q, H, W, accumulate)));
  PSG_CHECK_LAUNCH("psg_interleave2x2");
  return PSG_OK;
}

// out[r, c] = alpha * keep(r*C + c) * x[r, c] / (1 - p): same mask as the GEMM epilogue dropout with N == C.
int psg_dropout_scale(const void* x, long long ldx, void* out, long long ldo, long long rows, int C, float alpha,
                      unsigned long long seed, float drop_p, int dtype, void* stream) {
  PSG_CHECK_ARG(x && out && rows > 0 && C % 8 == 0 && ldx % 8 == 0 && ldo % 8 == 0, "psg_dropout_scale: bad args");
  const unsigned int thr = drop_p > 0.f ? (unsigned int)fmin((double)drop_p * 4294967296.0, 4294967295.0) : 0u;
  const float ks = drop_p > 0.f ? 1.f / (1.f - drop_p) : 1.f;
  DISPATCH_T(dtype, (dropout_scale_kernel<T><<<blocks_for(rows * (C / 8)), kThreads, 0, (cudaStream_t)stream>>>((const T*)x, ldx, (T*)out, ldo, rows, C, alpha, seed, thr, ks)));
  PSG_CHECK_LAUNCH("psg_dropout_scale");
  return PSG_OK;
}

int psg_timestep_embedding(const long long* t, const float* coeff, float* out, int B, int half, void* stream) {
  PSG_CHECK_ARG(B >= 0 && half > 0, "psg_timestep_embedding: bad sizes");
  if (B == 0) return PSG_OK;
  PSG_CHECK_ARG(t && coeff && out, "psg_timestep_embedding: null pointer");
  timestep_embedding_kernel<<<blocks_for((long long)B * half), kThreads, 0, (cudaStream_t)stream>>>(t, coeff, out, B, half);
  PSG_CHECK_LAUNCH("psg_timestep_embedding");
  return PSG_OK;
}

int psg_mean_pool(const float* x, float* out, int B, int L, int D, void* stream) {
  PSG_CHECK_ARG(B >= 0 && L > 0 && D > 0, "psg_mean_pool: bad sizes");
  if (B == 0) return PSG_OK;
  PSG_CHECK_ARG(x && out, "psg_mean_pool: null pointer");
  mean_pool_kernel<<<blocks_for((long long)B * D), kThreads, 0, (cudaStream_t)stream>>>(x, out, B, L, D);
  PSG_CHECK_LAUNCH("psg_mean_pool");
  return PSG_OK;
}

int psg_pack_conv_weight(const float* w, void* wp, void* wd, int Cout, int Cin, int kk, int Cin_p, int Cout_p, int dtype,
                         void* stream) {
  PSG_CHECK_ARG(w && (wp || wd) && Cout > 0 && Cin > 0 && kk > 0 && Cin_p >= Cin && Cout_p >= Cout, "psg_pack_conv_weight: bad args");
  PSG_CHECK_ARG(kk <= 25, "psg_pack_conv_weight: kernel area > 25 unsupported");
  dim3 grid((Cin + 31) / 32, (Cout + 31) / 32);
  PSG_CHECK_ARG(grid.y <= 65535, "psg_pack_conv_weight: Cout too large");
  const size_t smem = (size_t)32 * (32 * kk + 1) * sizeof(float);
  if (smem > 48 * 1024) {      // 4x4 / 5x5 kernels (the VAE encoder's stride-2 stem): opt in to the larger tile once
    static bool done = false;
    if (!done) {
      cudaFuncSetAttribute(pack_conv_kernel<float>, cudaFuncAttributeMaxDynamicSharedMemorySize, 32 * (32 * 25 + 1) * 4);
      cudaFuncSetAttribute(pack_conv_kernel<__nv_bfloat16>, cudaFuncAttributeMaxDynamicSharedMemorySize, 32 * (32 * 25 + 1) * 4);
      done = true;
    }
  }
  DISPATCH_T(dtype, (pack_conv_kernel<T><<<grid, 256, smem, (cudaStream_t)stream>>>(w, (T*)wp, (T*)wd, Cout, Cin, kk, Cin_p, Cout_p)));
  PSG_CHECK_LAUNCH("psg_pack_conv_weight");
  return PSG_OK;
}

// jobs: n x {src element offset, dst element offset, Cout, Cin, kk} (host array of long long[5]); channels % 64 == 0.
int psg_conv_weights_transpose(const void* src_base, void* dst_base, const long long* jobs, int n, void* stream) {
  PSG_CHECK_ARG(src_base && dst_base && jobs && n > 0 && n <= TransposeJobs::kMax, "psg_conv_weights_transpose: bad args (n=%d)", n);
  TransposeJobs tj;
  tj.n = n;
  int tiles = 0;
  for (int i = 0; i < n; ++i) {
    const long long* j = jobs + 5 * i;
    PSG_CHECK_ARG(j[2] % 64 == 0 && j[3] % 64 == 0 && j[4] > 0 && j[0] % 8 == 0 && j[1] % 8 == 0,
                  "psg_conv_weights_transpose: job %d: channels must be multiples of 64 and offsets of 8", i);
    tj.tile_begin[i] = tiles;
    tj.src[i] = j[0]; tj.dst[i] = j[1]; tj.cout[i] = (int)j[2]; tj.cin[i] = (int)j[3]; tj.kk[i] = (int)j[4];
    tiles += (int)(j[2] / 64 * (j[3] / 64) * j[4]);
  }
  tj.tile_begin[n] = tiles;
  conv_weight_transpose_kernel<<<tiles, 256, 0, (cudaStream_t)stream>>>((const __nv_bfloat16*)src_base, (__nv_bfloat16*)dst_base, tj);
  PSG_CHECK_LAUNCH("psg_conv_weights_transpose");
  return PSG_OK;
}

int psg_pack_linear_weight(const float* w, void* wk, void* wt, int N, int K, int dtype, void* stream) {
  PSG_CHECK_ARG(w && (wk || wt) && N > 0 && K > 0, "psg_pack_linear_weight: bad args");
  dim3 grid((K + 31) / 32, (N + 31) / 32);
  PSG_CHECK_ARG(grid.y <= 65535, "psg_pack_linear_weight: N too large");
  DISPATCH_T(dtype, (pack_linear_kernel<T><<<grid, 256, 0, (cudaStream_t)stream>>>(w, (T*)wk, (T*)wt, N, K)));
  PSG_CHECK_LAUNCH("psg_pack_linear_weight");
  return PSG_OK;
}

int psg_wgrad_finalize(const float* partial, int splits, long long split_stride, float* grad_oihw, int Cout, int Cin, int kk,
                       int Cin_p, int accumulate, void* stream) {
  PSG_CHECK_ARG(partial && grad_oihw && splits >= 1 && Cin_p >= Cin, "psg_wgrad_finalize: bad args");
  dim3 grid((Cin + 63) / 64, Cout);
  PSG_CHECK_ARG(Cout <= 65535, "psg_wgrad_finalize: Cout too large");
  wgrad_finalize_kernel<<<grid, 256, (size_t)kk * 65 * sizeof(float), (cudaStream_t)stream>>>(partial, splits, split_stride, grad_oihw, Cout, Cin, kk, Cin_p, accumulate);
  PSG_CHECK_LAUNCH("psg_wgrad_finalize");
  return PSG_OK;
}

int psg_sum_partials(const float* partial, int splits, long long split_stride, float* out, long long n, int accumulate, void* stream) {
  PSG_CHECK_ARG(partial && out && splits >= 1 && n > 0, "psg_sum_partials: bad args");
  sum_partials_kernel<<<blocks_for(n), kThreads, 0, (cudaStream_t)stream>>>(partial, splits, split_stride, out, n, accumulate);
  PSG_CHECK_LAUNCH("psg_sum_partials");
  return PSG_OK;
}

}  // extern "C"
