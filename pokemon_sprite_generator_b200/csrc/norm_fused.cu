// Single-pass GroupNorm (+SiLU) forward / backward for bf16 token-major activations: the (sample, channel-chunk) slab a
// block works on is staged ONCE in shared memory (cp.async, the whole slab in flight at once), so global traffic is
// the algorithmic minimum
//   forward : read x, write y                 (2 N elements)
//   backward: read x and dy, write dx         (3 N)
// instead of the two-pass scheme of norm_ops.cu (3 N / 5 N, the second pass hopefully from L2).  SURVEY.md §2.1 K4/K5.
//
// Reference ops replaced: nn.GroupNorm + F.silu (src/models/unet.py:79,89,115,127,156-157,214,231,397-398) and their
// autograd backward.
//
// Work decomposition: a chunk is CC channels, a multiple of lcm(channels-per-group, 8) (whole groups AND whole 16-byte
// vectors), so chunks are independent; one "unit" = (sample, chunk) is handled by TU = vpc * R threads (vpc vectors per
// pixel of the chunk, R pixel rows in flight), each thread owning one fixed vector column and the pixels row, row+R, ...
// A block carries U units.  The host picks (CC, R, U) per problem so that blocks are ~480 threads with few idle pixel
// slots (plan()).  All reductions are fixed-order (no atomics): results are run-to-run deterministic.
//
// The backward also emits, for free, the per-(sample, channel) sum over pixels of dx (analytically, from the same
// reductions): that is the gradient of the conv bias / broadcast conditioning that produced x (unet.py:116-124), which
// saves a separate column-sum pass over dx.
#include "norm_cluster.h"
#include "norm_stream.h"
#include "psg_common.cuh"

// 0: backward of tensors beyond the L2 by the streaming two-phase kernel (norm_stream.cu), cluster-split kernels where their plan
// applies, else the slab kernels; 1: slab kernels only; 2: streaming backward whenever its plan applies (tests); 3: as 0
// without the streaming backward (A/B)
static int g_gn_mode = 0;

namespace gnf {

constexpr int kMaxThreads = 480;   // divisible by 5 and 10 vector columns (channels/group 10, 20, 40, 80) and by 32
constexpr int kMaxIters = 8;
constexpr int kMaxNG = 32;         // groups per chunk
constexpr int kPS = 9;             // padded per-thread stride (floats) of the forward pair-sum scratch: conflict-free

struct FShape {
  int B, HW, C, G, cpg;
  int CC, vpc, NG;       // chunk: channels, vectors per pixel, groups
  int chunks;            // per sample
  int R, TU, U, iters;   // rows in flight per unit, threads per unit, units per block, pixel iterations
  int threads;
};

__host__ __device__ inline int gcd_i(int a, int b) { while (b) { int t = a % b; a = b; b = t; } return a; }

// Picks the chunk width and the row parallelism: maximise the fraction of live pixel slots HW / (iters * R), then the
// block size, then deeper per-thread pipelines (more bytes in flight per block), then narrow chunks.
inline int plan(FShape& s, int B, int HW, int C, int G) {
  if (B <= 0 || HW <= 0 || C <= 0 || G <= 0 || C % G != 0 || C % 8 != 0) return -1;
  s.B = B; s.HW = HW; s.C = C; s.G = G; s.cpg = C / G;
  if (s.cpg % 2 != 0) return -1;
  const int cc0 = s.cpg / gcd_i(s.cpg, 8) * 8;
  if (C % cc0 != 0) return -1;
  const int nbase = C / cc0;
  double best = -1.0;
  for (int m = 1; m <= nbase; ++m) {
    if (nbase % m) continue;
    const int cc = cc0 * m, vpc = cc / 8, ng = cc / s.cpg;
    if (vpc > kMaxThreads || ng > kMaxNG) break;
    const int r_unit = 32 / gcd_i(vpc, 32);           // TU = vpc * R must be a whole number of warps
    for (int r = r_unit; r * vpc <= kMaxThreads; r += r_unit) {
      const int iters = (HW + r - 1) / r;
      if (iters <= kMaxIters) {
        const int tu = vpc * r;
        int u = kMaxThreads / tu;
        if (u > B * (C / cc)) u = B * (C / cc);
        const double eff = (double)HW / ((double)iters * r);
        const double score = eff * 1000.0 + (double)(tu * u) / kMaxThreads * 10.0 + iters * 0.1 - m * 0.001;   // more bytes in flight per block
        if (score > best) {
          best = score;
          s.CC = cc; s.vpc = vpc; s.NG = ng; s.chunks = C / cc; s.R = r; s.TU = tu; s.U = u; s.iters = iters; s.threads = tu * u;
        }
      }
      if (r >= HW) break;
    }
  }
  return best < 0.0 ? -1 : 0;
}

__device__ __forceinline__ void unpack8(const uint4& r, float (&v)[8]) {
  const __nv_bfloat162* h = reinterpret_cast<const __nv_bfloat162*>(&r);
#pragma unroll
  for (int i = 0; i < 4; ++i) { float2 f = __bfloat1622float2(h[i]); v[2 * i] = f.x; v[2 * i + 1] = f.y; }
}
__device__ __forceinline__ uint4 pack8(const float (&v)[8]) {
  uint4 r;
  __nv_bfloat162* h = reinterpret_cast<__nv_bfloat162*>(&r);
#pragma unroll
  for (int i = 0; i < 4; ++i) h[i] = __floats2bfloat162_rn(v[2 * i], v[2 * i + 1]);
  return r;
}
// 16-byte asynchronous global->shared copy (LDGSTS); src_bytes == 0 zero-fills the slot without touching memory.
__device__ __forceinline__ void cp_async16(void* smem_dst, const void* gmem_src, int src_bytes) {
  const uint32_t d = (uint32_t)__cvta_generic_to_shared(smem_dst);
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(d), "l"(gmem_src), "r"(src_bytes) : "memory");
}
__device__ __forceinline__ void cp_async_wait_all() {
  asm volatile("cp.async.commit_group;\n\tcp.async.wait_group 0;" ::: "memory");
}
// sigmoid through the single-MUFU tanh approximation (rel. error ~2^-11: below the bf16 rounding of the results)
__device__ __forceinline__ float fast_sigmoid(float x) {
  float t;
  asm("tanh.approx.f32 %0, %1;" : "=f"(t) : "f"(0.5f * x));
  return fmaf(0.5f, t, 0.5f);
}
__device__ __forceinline__ float fast_silu(float x) { return x * fast_sigmoid(x); }
__device__ __forceinline__ float fast_silu_grad(float x) {
  const float s = fast_sigmoid(x);
  return s * fmaf(x, 1.f - s, 1.f);
}

// ---------------------------------------------------------------------------------------------------------------------
// forward
// ---------------------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(kMaxThreads, 2) gn_fused_fwd_kernel(const __nv_bfloat16* __restrict__ x, long long ld,
                                                                     __nv_bfloat16* __restrict__ y, long long ldy,
                                                                     const float* __restrict__ gamma,
                                                                     const float* __restrict__ beta, FShape s,
                                                                     float* __restrict__ stats, float eps, int act) {
  extern __shared__ uint4 slab[];                                             // [iters][threads] staged x vectors
  const int T = s.threads, tid = threadIdx.x;
  float* pair = reinterpret_cast<float*>(slab + (size_t)s.iters * T);         // [threads][kPS] pair sums / squares
  float* st = pair + (size_t)T * kPS;                                         // [U][NG][2] mean, rstd
  const int u = tid / s.TU, tu = tid - u * s.TU;
  const int unit = blockIdx.x * s.U + u;
  const bool active = unit < s.B * s.chunks;
  const int b = active ? unit / s.chunks : 0;
  const int ch = active ? unit - b * s.chunks : 0;
  const int v = tu % s.vpc, row = tu / s.vpc;
  const int c0 = ch * s.CC + v * 8;
  const __nv_bfloat16* xb = x + (long long)b * s.HW * ld + c0;
  uint4* mine = slab + tid;

  // the whole slab is requested at once (every 16-byte piece in flight before the first wait); each thread later reads
  // back only the slots it copied itself, so no block barrier is needed on the data
  for (int k = 0; k < s.iters; ++k) {
    const int p = row + k * s.R;
    const bool ok = active && p < s.HW;
    cp_async16(mine + (size_t)k * T, ok ? xb + (long long)p * ld : x, ok ? 16 : 0);
  }
  float ga[8], be[8];      // affine parameters: fetched while the slab is in flight
  {
    const float4 g0 = __ldg(reinterpret_cast<const float4*>(gamma + c0)), g1 = __ldg(reinterpret_cast<const float4*>(gamma + c0 + 4));
    const float4 b0 = __ldg(reinterpret_cast<const float4*>(beta + c0)), b1 = __ldg(reinterpret_cast<const float4*>(beta + c0 + 4));
    ga[0] = g0.x; ga[1] = g0.y; ga[2] = g0.z; ga[3] = g0.w; ga[4] = g1.x; ga[5] = g1.y; ga[6] = g1.z; ga[7] = g1.w;
    be[0] = b0.x; be[1] = b0.y; be[2] = b0.z; be[3] = b0.w; be[4] = b1.x; be[5] = b1.y; be[6] = b1.z; be[7] = b1.w;
  }
  cp_async_wait_all();
  float ps[4] = {0.f, 0.f, 0.f, 0.f}, pq[4] = {0.f, 0.f, 0.f, 0.f};
  for (int k = 0; k < s.iters; ++k) {
    float f[8];
    unpack8(mine[(size_t)k * T], f);
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      ps[j] += f[2 * j] + f[2 * j + 1];
      pq[j] += f[2 * j] * f[2 * j] + f[2 * j + 1] * f[2 * j + 1];
    }
  }
  {
    float* m = pair + (size_t)tid * kPS;
#pragma unroll
    for (int j = 0; j < 4; ++j) { m[j] = ps[j]; m[4 + j] = pq[j]; }
  }
  __syncthreads();
  // one warp per group: lanes walk the unit's rows, then a shuffle fold (a channel pair never straddles a group)
  {
    const int wpu = s.TU >> 5, wu = tu >> 5, lane = tid & 31;
    const int ppg = s.cpg >> 1;                         // pairs per group
    for (int g = wu; g < s.NG; g += wpu) {
      float su = 0.f, sq = 0.f;
      for (int r = lane; r < s.R; r += 32) {
        const float* base = pair + ((size_t)u * s.TU + (size_t)r * s.vpc) * kPS;
        for (int pi = g * ppg; pi < (g + 1) * ppg; ++pi) {
          const float* src = base + (pi >> 2) * kPS + (pi & 3);
          su += src[0];
          sq += src[4];
        }
      }
      su = psg_warp_sum(su);
      sq = psg_warp_sum(sq);
      if (lane == 0) {
        const float m = (float)s.cpg * (float)s.HW;
        const float mean = su / m;
        const float var = fmaxf(sq / m - mean * mean, 0.f);
        const float rstd = rsqrtf(var + eps);
        st[(u * s.NG + g) * 2] = mean;
        st[(u * s.NG + g) * 2 + 1] = rstd;
        if (active) {
          float* o = stats + ((long long)b * s.G + ch * s.NG + g) * 2;
          o[0] = mean;
          o[1] = rstd;
        }
      }
    }
  }
  __syncthreads();
  if (!active) return;
  float sc[8], sh[8];      // y = x * sc + sh
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    const int gi = (v * 8 + j) / s.cpg;
    const float mean = st[(u * s.NG + gi) * 2], rstd = st[(u * s.NG + gi) * 2 + 1];
    sc[j] = rstd * ga[j];
    sh[j] = be[j] - mean * sc[j];
  }
  __nv_bfloat16* yb = y + (long long)b * s.HW * ldy + c0;
  for (int k = 0; k < s.iters; ++k) {
    const int p = row + k * s.R;
    if (p >= s.HW) break;
    float f[8];
    unpack8(mine[(size_t)k * T], f);
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const float n = fmaf(f[j], sc[j], sh[j]);
      f[j] = act ? fast_silu(n) : n;
    }
    *reinterpret_cast<uint4*>(yb + (long long)p * ldy) = pack8(f);
  }
}

// ---------------------------------------------------------------------------------------------------------------------
// backward
//   xh = (x - mean) * rstd ; n = xh*gamma + beta ; dn = dy * act'(n)
//   s1[c] = sum_pix dn ; s2[c] = sum_pix dn*xh              (-> dbeta, dgamma after the sum over samples)
//   s3[c] = sum_pix xh
//   A[g] = sum_{c in g} gamma[c]*s1[c] ; Bs[g] = sum_{c in g} gamma[c]*s2[c]
//   dx = rstd * (dn*gamma - (A + xh*Bs)/m)
//   sum_pix dx[c] = rstd * (gamma[c]*s1[c] - (HW*A + s3[c]*Bs)/m)
// partial[b][c] = {s1, s2, sum_pix dx}
// ---------------------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(kMaxThreads) gn_fused_bwd_kernel(const __nv_bfloat16* __restrict__ dy, long long lddy,
                                                                  const __nv_bfloat16* __restrict__ x, long long ld,
                                                                  __nv_bfloat16* __restrict__ dx, long long lddx,
                                                                  const float* __restrict__ gamma, const float* __restrict__ beta,
                                                                  const float* __restrict__ stats, FShape s,
                                                                  float* __restrict__ partial, float* __restrict__ colsum_out,
                                                                  long long ld_colsum, int act, int accumulate) {
  extern __shared__ uint4 slab[];                       // [2][iters][threads] staged x and dy vectors
  const int T = s.threads, tid = threadIdx.x;
  uint4* slab_x = slab + tid;
  uint4* slab_d = slab + (size_t)s.iters * T + tid;
  float* acc = reinterpret_cast<float*>(slab + (size_t)2 * s.iters * T);   // [threads][24]
  float* chan = acc + (size_t)T * 24;                                       // [U][CC][3]
  float* grp = chan + (size_t)s.U * s.CC * 3;                               // [U][NG][2]
  const int u = tid / s.TU, tu = tid - u * s.TU;
  const int unit = blockIdx.x * s.U + u;
  const bool active = unit < s.B * s.chunks;
  const int b = active ? unit / s.chunks : 0;
  const int ch = active ? unit - b * s.chunks : 0;
  const int v = tu % s.vpc, row = tu / s.vpc;
  const int c0 = ch * s.CC + v * 8;
  const __nv_bfloat16* xb = x + (long long)b * s.HW * ld + c0;
  const __nv_bfloat16* db = dy + (long long)b * s.HW * lddy + c0;

  for (int k = 0; k < s.iters; ++k) {
    const int p = row + k * s.R;
    const bool ok = active && p < s.HW;
    cp_async16(slab_x + (size_t)k * T, ok ? xb + (long long)p * ld : x, ok ? 16 : 0);
    cp_async16(slab_d + (size_t)k * T, ok ? db + (long long)p * lddy : dy, ok ? 16 : 0);
  }
  float ga[8], be[8], mu[8], rs[8];    // fetched while the slab is in flight
  {
    const float4 g0 = __ldg(reinterpret_cast<const float4*>(gamma + c0)), g1 = __ldg(reinterpret_cast<const float4*>(gamma + c0 + 4));
    const float4 b0 = __ldg(reinterpret_cast<const float4*>(beta + c0)), b1 = __ldg(reinterpret_cast<const float4*>(beta + c0 + 4));
    ga[0] = g0.x; ga[1] = g0.y; ga[2] = g0.z; ga[3] = g0.w; ga[4] = g1.x; ga[5] = g1.y; ga[6] = g1.z; ga[7] = g1.w;
    be[0] = b0.x; be[1] = b0.y; be[2] = b0.z; be[3] = b0.w; be[4] = b1.x; be[5] = b1.y; be[6] = b1.z; be[7] = b1.w;
  }
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    const int g = (c0 + j) / s.cpg;
    const float2 mr = __ldg(reinterpret_cast<const float2*>(stats + ((long long)b * s.G + g) * 2));
    mu[j] = mr.x;
    rs[j] = mr.y;
  }
  cp_async_wait_all();
  float a1[8], a2[8], a3[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) { a1[j] = 0.f; a2[j] = 0.f; a3[j] = 0.f; }
  for (int k = 0; k < s.iters; ++k) {
    const int p = row + k * s.R;
    const float live = (active && p < s.HW) ? 1.f : 0.f;
    float fx[8], fd[8];
    unpack8(slab_x[(size_t)k * T], fx);
    unpack8(slab_d[(size_t)k * T], fd);
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const float xh = (fx[j] - mu[j]) * rs[j];
      const float dn = act ? fd[j] * fast_silu_grad(fmaf(xh, ga[j], be[j])) : fd[j];   // fd == 0 on padded slots
      a1[j] += dn;
      a2[j] = fmaf(dn, xh, a2[j]);
      a3[j] = fmaf(live, xh, a3[j]);
      fd[j] = dn;
    }
    if (act) slab_d[(size_t)k * T] = pack8(fd);     // keep dn (bf16) for the second pass instead of re-evaluating SiLU'
  }
  {
    float4* mine = reinterpret_cast<float4*>(acc + (size_t)tid * 24);
    mine[0] = make_float4(a1[0], a1[1], a1[2], a1[3]);
    mine[1] = make_float4(a1[4], a1[5], a1[6], a1[7]);
    mine[2] = make_float4(a2[0], a2[1], a2[2], a2[3]);
    mine[3] = make_float4(a2[4], a2[5], a2[6], a2[7]);
    mine[4] = make_float4(a3[0], a3[1], a3[2], a3[3]);
    mine[5] = make_float4(a3[4], a3[5], a3[6], a3[7]);
  }
  __syncthreads();
  // per-channel fold over the R rows of the unit, fixed order: thread (stat, c) of the unit
  for (int idx = tu; idx < 3 * s.CC; idx += s.TU) {
    const int stat = idx / s.CC, c = idx - stat * s.CC;
    const float* src = acc + ((size_t)u * s.TU + (c >> 3)) * 24 + stat * 8 + (c & 7);
    float t0 = 0.f, t1 = 0.f, t2 = 0.f, t3 = 0.f;
    int r = 0;
    for (; r + 3 < s.R; r += 4) {
      t0 += src[(size_t)(r + 0) * s.vpc * 24];
      t1 += src[(size_t)(r + 1) * s.vpc * 24];
      t2 += src[(size_t)(r + 2) * s.vpc * 24];
      t3 += src[(size_t)(r + 3) * s.vpc * 24];
    }
    for (; r < s.R; ++r) t0 += src[(size_t)r * s.vpc * 24];
    chan[((size_t)u * s.CC + c) * 3 + stat] = (t0 + t1) + (t2 + t3);
  }
  __syncthreads();
  for (int idx = tu; idx < 2 * s.NG; idx += s.TU) {
    const int stat = idx & 1, g = idx >> 1;
    float t = 0.f;
    for (int c = g * s.cpg; c < (g + 1) * s.cpg; ++c) t = fmaf(__ldg(gamma + ch * s.CC + c), chan[((size_t)u * s.CC + c) * 3 + stat], t);
    grp[(u * s.NG + g) * 2 + stat] = t;
  }
  __syncthreads();
  if (!active) return;
  const float inv_m = 1.f / ((float)s.cpg * (float)s.HW);
  for (int c = tu; c < s.CC; c += s.TU) {
    const int g = c / s.cpg, cg = ch * s.CC + c;
    const float s1 = chan[((size_t)u * s.CC + c) * 3], s2 = chan[((size_t)u * s.CC + c) * 3 + 1];
    const float s3 = chan[((size_t)u * s.CC + c) * 3 + 2];
    const float A = grp[(u * s.NG + g) * 2], Bs = grp[(u * s.NG + g) * 2 + 1];
    const float rstd = __ldg(stats + ((long long)b * s.G + cg / s.cpg) * 2 + 1);
    const float cs = rstd * (__ldg(gamma + cg) * s1 - ((float)s.HW * A + s3 * Bs) * inv_m);
    float* o = partial + ((long long)b * s.C + cg) * 3;
    o[0] = s1;
    o[1] = s2;
    o[2] = cs;
    if (colsum_out) colsum_out[(long long)b * ld_colsum + cg] = cs;
  }
  // dx = fd * k1 - k2 - xh * k3  with  k1 = rstd*gamma, k2 = rstd*A/m, k3 = rstd*Bs/m
  float k1[8], k2[8], k3[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    const int gi = (v * 8 + j) / s.cpg;
    k1[j] = rs[j] * ga[j];
    k2[j] = rs[j] * grp[(u * s.NG + gi) * 2] * inv_m;
    k3[j] = rs[j] * grp[(u * s.NG + gi) * 2 + 1] * inv_m;
  }
  __nv_bfloat16* ob = dx + (long long)b * s.HW * lddx + c0;
  for (int k = 0; k < s.iters; ++k) {
    const int p = row + k * s.R;
    if (p >= s.HW) break;
    float fx[8], fd[8], fo[8];
    unpack8(slab_x[(size_t)k * T], fx);
    unpack8(slab_d[(size_t)k * T], fd);
    if (accumulate) unpack8(*reinterpret_cast<const uint4*>(ob + (long long)p * lddx), fo);
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const float xh = (fx[j] - mu[j]) * rs[j];
      const float d = fmaf(fd[j], k1[j], -fmaf(xh, k3[j], k2[j]));      // fd holds dn
      fo[j] = accumulate ? fo[j] + d : d;
    }
    *reinterpret_cast<uint4*>(ob + (long long)p * lddx) = pack8(fo);
  }
}

// dbeta[c] = sum_b partial[b][c][0]; dgamma[c] = sum_b partial[b][c][1]; bias_total[c] = sum_b partial[b][c][2].
// 8 channels x 64 sample-lanes per block (C / 8 blocks: the 10-block grid of the earlier 32 x 16 shape took 7 us per norm
// at C = 320, latency bound: profiles/r02_ncu_groupnorm_stream_bwd.md), every lane's rows in flight together, fixed-order fold.
constexpr int kPgCh = 8, kPgLanes = 64;
__global__ void __launch_bounds__(kPgCh * kPgLanes) gn_fused_param_grad_kernel(const float* __restrict__ partial, int B, int C,
                                                                               float* __restrict__ dgamma, float* __restrict__ dbeta,
                                                                               float* __restrict__ bias_total, int accumulate) {
  __shared__ float sh[kPgLanes][kPgCh * 3];
  const int tx = threadIdx.x % (kPgCh * 3), ty = threadIdx.x / (kPgCh * 3);     // 24 consecutive floats of a row x 21 row lanes
  constexpr int kRowLanes = kPgCh * kPgLanes / (kPgCh * 3);                       // 21 (8 threads idle)
  const int c0 = blockIdx.x * kPgCh;
  const int ncol = min(kPgCh, C - c0) * 3;
  float t = 0.f;
  if (ty < kRowLanes && tx < ncol) {
    const float* p = partial + ((long long)ty * C + c0) * 3 + tx;
    const long long step = (long long)kRowLanes * C * 3;
    int r = ty;
    for (; r + 7 * kRowLanes < B; r += 8 * kRowLanes, p += 8 * step) {      // eight rows' loads in flight per thread (latency bound)
      float a[8];
#pragma unroll
      for (int k = 0; k < 8; ++k) a[k] = p[k * step];
#pragma unroll
      for (int k = 0; k < 8; ++k) t += a[k];
    }
    for (; r + 3 * kRowLanes < B; r += 4 * kRowLanes, p += 4 * step) {
      const float a = p[0], b = p[step], c = p[2 * step], d = p[3 * step];
      t += a; t += b; t += c; t += d;
    }
    for (; r < B; r += kRowLanes, p += step) t += p[0];
  }
  if (ty < kRowLanes) sh[ty][tx] = t;
  __syncthreads();
  if (threadIdx.x < ncol) {
    float u = 0.f;
#pragma unroll
    for (int k = 0; k < kRowLanes; ++k) u += sh[k][threadIdx.x];
    const int c = c0 + threadIdx.x / 3, stat = threadIdx.x % 3;
    if (stat == 0) dbeta[c] = accumulate ? dbeta[c] + u : u;
    else if (stat == 1) dgamma[c] = accumulate ? dgamma[c] + u : u;
    else if (bias_total) bias_total[c] = u;
  }
}
#define PSG_GN_PARAM_GRAD_GRID(C) (((C) + gnf::kPgCh - 1) / gnf::kPgCh), (gnf::kPgCh * gnf::kPgLanes)

static size_t fwd_smem(const FShape& s) {
  return (size_t)s.iters * s.threads * 16 + ((size_t)s.threads * kPS + (size_t)s.U * s.NG * 2) * sizeof(float);
}
static size_t bwd_smem(const FShape& s) {
  return (size_t)2 * s.iters * s.threads * 16 +
         ((size_t)s.threads * 24 + (size_t)s.U * s.CC * 3 + (size_t)s.U * s.NG * 2) * sizeof(float);
}
constexpr size_t kSmemLimit = 220 * 1024;

}  // namespace gnf

extern "C" {

// 1 if the single-pass kernels handle this problem (bf16, slab fits in shared memory); else the two-pass kernels run.
int psg_groupnorm_fused_mode(int mode) {
  const int prev = g_gn_mode;
  if (mode >= 0 && mode <= 3) g_gn_mode = mode;
  return prev;
}

// tunables of the cluster-split kernels (measurement hook): which = 0 fwd threads, 1 bwd threads, 2 bytes of one tensor per
// CTA, 3 largest cluster size; value <= 0 only reads.  Returns the previous value.
int psg_groupnorm_cluster_tune(int which, int value) { return gnc_tune(which, value); }

// out8 = {CC, cluster size, rows per CTA, R, TU, U, iters, smem bytes} of the cluster-split kernels (norm_cluster.cu)
int psg_groupnorm_cluster_plan(int B, int HW, int C, int G, int bwd, int* out) {
  PSG_CHECK_ARG(out && gnc_plan(B, HW, C, G, bwd, out) == 0, "psg_groupnorm_cluster_plan: unsupported shape B=%d HW=%d C=%d G=%d", B, HW, C, G);
  return PSG_OK;
}

int psg_groupnorm_fused_ok(int B, int HW, int C, int G, int dtype) {
  if (dtype != PSG_DTYPE_BF16) return 0;
  if (g_gn_mode != 1 && gnc_supported(B, HW, C, G, 0) && gnc_supported(B, HW, C, G, 1)) return 1;
  gnf::FShape s;
  if (gnf::plan(s, B, HW, C, G) != 0) return 0;
  return gnf::bwd_smem(s) <= gnf::kSmemLimit ? 1 : 0;
}

// Reports the decomposition plan() picks: out = {CC, R, TU, U, iters, threads, fwd smem bytes, bwd smem bytes}.
int psg_groupnorm_fused_plan(int B, int HW, int C, int G, int* out) {
  gnf::FShape s;
  PSG_CHECK_ARG(out && gnf::plan(s, B, HW, C, G) == 0, "psg_groupnorm_fused_plan: unsupported shape B=%d HW=%d C=%d G=%d", B, HW, C, G);
  out[0] = s.CC; out[1] = s.R; out[2] = s.TU; out[3] = s.U; out[4] = s.iters; out[5] = s.threads;
  out[6] = (int)gnf::fwd_smem(s); out[7] = (int)gnf::bwd_smem(s);
  return PSG_OK;
}

int psg_groupnorm_fused_fwd(const void* x, long long ld_x, void* y, long long ld_y, const float* gamma, const float* beta,
                            float* stats, int B, int HW, int C, int G, float eps, int act, void* stream) {
  using namespace gnf;
  PSG_CHECK_ARG(x && y && gamma && beta && stats, "psg_groupnorm_fused_fwd: null pointer");
  PSG_CHECK_ARG(ld_x % 8 == 0 && ld_y % 8 == 0 && ((uintptr_t)x % 16 == 0) && ((uintptr_t)y % 16 == 0) &&
                    ((uintptr_t)gamma % 16 == 0) && ((uintptr_t)beta % 16 == 0),
                "psg_groupnorm_fused_fwd: pitches/pointers must be 16B aligned");
  if (g_gn_mode != 1) {
    const int rc = gnc_fwd(x, ld_x, y, ld_y, gamma, beta, stats, B, HW, C, G, eps, act, (cudaStream_t)stream);
    if (rc != PSG_ERR_UNSUPPORTED) return rc;
  }
  FShape s;
  PSG_CHECK_ARG(plan(s, B, HW, C, G) == 0 && fwd_smem(s) <= kSmemLimit,
                "psg_groupnorm_fused_fwd: unsupported shape B=%d HW=%d C=%d G=%d", B, HW, C, G);
  static bool configured = false;
  if (!configured) {
    cudaError_t e = cudaFuncSetAttribute(gn_fused_fwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kSmemLimit);
    if (e != cudaSuccess) { psg_set_error("psg_groupnorm_fused_fwd: cudaFuncSetAttribute: %s", cudaGetErrorString(e)); return PSG_ERR_CUDA; }
    configured = true;
  }
  const int units = B * s.chunks;
  const int grid = (units + s.U - 1) / s.U;
  gn_fused_fwd_kernel<<<grid, s.threads, fwd_smem(s), (cudaStream_t)stream>>>((const __nv_bfloat16*)x, ld_x, (__nv_bfloat16*)y, ld_y,
                                                                              gamma, beta, s, stats, eps, act);
  PSG_CHECK_LAUNCH("psg_groupnorm_fused_fwd");
  return PSG_OK;
}

// workspace floats >= B*C*3.  dx_colsum [B, ld_colsum] and bias_total [C] are optional (null): the per-(sample, channel)
// and per-channel sums over pixels of dx (written, not accumulated; valid only when accumulate_dx == 0).
// workspace_floats: size of `workspace`; >= psg_groupnorm_bwd_workspace_floats(B, HW, C, G) enables the streaming two-phase
// backward for tensors beyond the L2 (norm_stream.cu); with B*C*3 floats the single-pass kernels run.
int psg_groupnorm_fused_bwd_ws(const void* dy, long long ld_dy, const void* x, long long ld_x, void* dx, long long ld_dx,
                               const float* gamma, const float* beta, const float* stats, float* dgamma, float* dbeta,
                               float* workspace, long long workspace_floats, float* dx_colsum, long long ld_colsum, float* bias_total,
                               int B, int HW, int C, int G, int act, int accumulate_dx, int accumulate_params, void* stream) {
  using namespace gnf;
  PSG_CHECK_ARG(workspace_floats >= (long long)B * C * 3, "psg_groupnorm_fused_bwd: workspace of %lld floats < B*C*3", workspace_floats);
  PSG_CHECK_ARG(dy && x && dx && gamma && beta && stats && dgamma && dbeta && workspace, "psg_groupnorm_fused_bwd: null pointer");
  PSG_CHECK_ARG(!(accumulate_dx && (dx_colsum || bias_total)), "psg_groupnorm_fused_bwd: column sums of dx need accumulate_dx == 0");
  PSG_CHECK_ARG(ld_x % 8 == 0 && ld_dy % 8 == 0 && ld_dx % 8 == 0 && ((uintptr_t)x % 16 == 0) && ((uintptr_t)dy % 16 == 0) &&
                    ((uintptr_t)dx % 16 == 0) && ((uintptr_t)gamma % 16 == 0) && ((uintptr_t)beta % 16 == 0),
                "psg_groupnorm_fused_bwd: pitches/pointers must be 16B aligned");
  if ((g_gn_mode == 0 && gns_wants(B, HW, C)) || g_gn_mode == 2) {
    const int rc = gns_bwd(dy, ld_dy, x, ld_x, dx, ld_dx, gamma, beta, stats, workspace, workspace_floats, dx_colsum, ld_colsum, B, HW, C, G,
                           act, accumulate_dx, (cudaStream_t)stream);
    if (rc == PSG_OK) {
      gn_fused_param_grad_kernel<<<PSG_GN_PARAM_GRAD_GRID(C), 0, (cudaStream_t)stream>>>(workspace, B, C, dgamma, dbeta, bias_total, accumulate_params);
      PSG_CHECK_LAUNCH("psg_groupnorm_fused_bwd");
      return PSG_OK;
    }
    if (rc != PSG_ERR_UNSUPPORTED) return rc;
  }
  if (g_gn_mode != 1) {
    const int rc = gnc_bwd(dy, ld_dy, x, ld_x, dx, ld_dx, gamma, beta, stats, workspace, dx_colsum, ld_colsum, B, HW, C, G, act,
                           accumulate_dx, (cudaStream_t)stream);
    if (rc == PSG_OK) {
      gn_fused_param_grad_kernel<<<PSG_GN_PARAM_GRAD_GRID(C), 0, (cudaStream_t)stream>>>(workspace, B, C, dgamma, dbeta, bias_total, accumulate_params);
      PSG_CHECK_LAUNCH("psg_groupnorm_fused_bwd");
      return PSG_OK;
    }
    if (rc != PSG_ERR_UNSUPPORTED) return rc;
  }
  FShape s;
  PSG_CHECK_ARG(plan(s, B, HW, C, G) == 0 && bwd_smem(s) <= kSmemLimit,
                "psg_groupnorm_fused_bwd: unsupported shape B=%d HW=%d C=%d G=%d", B, HW, C, G);
  static bool configured = false;
  if (!configured) {
    cudaError_t e = cudaFuncSetAttribute(gn_fused_bwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kSmemLimit);
    if (e != cudaSuccess) { psg_set_error("psg_groupnorm_fused_bwd: cudaFuncSetAttribute: %s", cudaGetErrorString(e)); return PSG_ERR_CUDA; }
    configured = true;
  }
  const int units = B * s.chunks;
  const int grid = (units + s.U - 1) / s.U;
  cudaStream_t st = (cudaStream_t)stream;
  gn_fused_bwd_kernel<<<grid, s.threads, bwd_smem(s), st>>>((const __nv_bfloat16*)dy, ld_dy, (const __nv_bfloat16*)x, ld_x,
                                                            (__nv_bfloat16*)dx, ld_dx, gamma, beta, stats, s, workspace, dx_colsum,
                                                            ld_colsum, act, accumulate_dx);
  gn_fused_param_grad_kernel<<<PSG_GN_PARAM_GRAD_GRID(C), 0, st>>>(workspace, B, C, dgamma, dbeta, bias_total, accumulate_params);
  PSG_CHECK_LAUNCH("psg_groupnorm_fused_bwd");
  g_psg_launch_count += 1;  // two kernels
  return PSG_OK;
}


int psg_groupnorm_fused_bwd(const void* dy, long long ld_dy, const void* x, long long ld_x, void* dx, long long ld_dx,
                            const float* gamma, const float* beta, const float* stats, float* dgamma, float* dbeta,
                            float* workspace, float* dx_colsum, long long ld_colsum, float* bias_total, int B, int HW, int C, int G,
                            int act, int accumulate_dx, int accumulate_params, void* stream) {
  return psg_groupnorm_fused_bwd_ws(dy, ld_dy, x, ld_x, dx, ld_dx, gamma, beta, stats, dgamma, dbeta, workspace, (long long)B * C * 3,
                                    dx_colsum, ld_colsum, bias_total, B, HW, C, G, act, accumulate_dx, accumulate_params, stream);
}

// floats of workspace that let psg_groupnorm_fused_bwd_ws pick any of its kernels for this shape (>= B*C*3)
long long psg_groupnorm_bwd_workspace_floats(int B, int HW, int C, int G) {
  const long long base = (long long)B * C * 3, st = gns_workspace_floats(B, HW, C, G);
  return st > base ? st : base;
}
// streaming backward: tunables (0 bytes of x + dy per L2-resident sample group, 1 pixel rows per chunk, 2 smallest tensor in bytes
// of x that the auto mode routes to it; value < 0 only reads; returns the previous value) and plan
// (out8 = {vectors per row, row lanes, threads, rows per chunk, chunks per sample, samples per group, smem bytes, grid})
long long psg_groupnorm_stream_tune(int which, long long value) { return gns_tune(which, value); }
int psg_groupnorm_stream_plan(int B, int HW, int C, int G, int* out) {
  PSG_CHECK_ARG(out && gns_plan(B, HW, C, G, out) == 0, "psg_groupnorm_stream_plan: unsupported shape B=%d HW=%d C=%d G=%d", B, HW, C, G);
  return PSG_OK;
}
// Test hook: 1 if an apply CTA's bounded wait for its sample's moments expired since the last call (synchronises the device).
int psg_groupnorm_timeout_flag() { return gns_timeout_flag(); }

}  // extern "C"
