// Single-pass GroupNorm (+SiLU) forward / backward, pixel-split across a thread-block cluster (bf16, token-major).
//
// Reference ops replaced: nn.GroupNorm + F.silu (src/models/unet.py:79,89,115,127,156-157,214,231,397-398) and their
// autograd backward.  SURVEY.md §2.1 K4/K5; bound by HBM (forward 2 N, backward 3 N elements of traffic).
//
// Why a second single-pass kernel next to norm_fused.cu.  ncu on the slab kernels (profiles/r01_ncu_groupnorm_v1.md)
// shows them ISSUE bound, not memory bound: 26 (forward) / 52 (backward) thread-instructions per element at 55-65 % issue
// utilisation, most of it per-thread set-up (runtime integer divisions, 64-bit index multiplies, un-unrolled loops)
// amortised over only 32-64 elements, plus 80-byte row pieces (1.3x DRAM over-fetch).  Here
//   * the kernels are specialised at compile time for the U-Net's shapes: a unit is (sample, 160 channels = 320 B rows,
//     i.e. whole 64 B sectors) and channels-per-group is a template parameter (10, 20, 40, 80: C = 320 ... 2560 with 32
//     groups), so every division is by a constant and every loop walks pointers;
//   * the arithmetic runs on packed fp32 pairs (FFMA2/FADD2/FMUL2, sm_100): affine maps folded into one FMA per use,
//     SiLU and its derivative from one tanh.approx per element;
//   * the unit's pixels are split across the S CTAs of a cluster (<= 32 KB of x per CTA), which exchange their partial
//     sums through distributed shared memory in rank order, so several CTAs are resident per SM and their load, reduce
//     and store phases overlap.
// Every reduction is fixed-order: results are run-to-run bit-identical.  Other shapes fall back to norm_fused.cu.
#include <cooperative_groups.h>

#include <cstdlib>

#include "norm_cluster.h"
#include "psg_common.cuh"

namespace cg = cooperative_groups;

namespace gnc {

constexpr int kMaxThreads = 512;
constexpr int kMaxIters = 16;
constexpr size_t kSmemLimit = 200 * 1024;

// tunables (psg_groupnorm_cluster_tune): {fwd threads, bwd threads, fwd bytes of x per CTA, largest cluster, 16-byte
// vectors per pixel row of a unit (20: 160 channels = 320 B rows; 10: 80 channels = 160 B rows; 0: by shape), bwd bytes
// of x per CTA (0: by shape), L2 prefetch of the rows of the CTA that will run one residency later (0: off, 1: distance
// from the occupancy estimate below, > 1: that many resident CTAs)}.  Defaults from tools/sweep_gn.py on B200
// (profiles/r01_bench_groupnorm_v2.txt); the environment variable PSG_GN_PREFETCH overrides entry 6 (A/B runs).
static int g_tune[7] = {160, 160, 32 * 1024, 8, 0, 0, 296};
static bool g_tune_env_read = false;

struct Shape {
  int B, HW, C, G, cpg;
  int vpc, chunks;             // 16-byte vectors per pixel row of a unit (unit = 8 vpc channels), units per sample
  int S, rp;                   // cluster size, pixel rows per CTA
  int R, TU, U, iters;         // rows in flight per unit, threads per unit, units per block, pixel iterations
  int threads;
  int pf;                      // L2 prefetch distance in units (the same rank of the cluster `pf` units ahead); 0 = none
  int strict;                  // 1: every thread arrives with .release (A/B switch, env PSG_GN_STRICT_BARRIER=1)
  int pf_dx;                   // backward with accumulate: ask L2 for this CTA's own dx rows while x and dy are in flight
};

static int plan(Shape& s, int B, int HW, int C, int G, int bwd) {
  const int kVPC = g_tune[4] ? g_tune[4] : (HW >= 128 ? 10 : 20), kCC = kVPC * 8;
  const int slab_target = !bwd ? g_tune[2] : g_tune[5] ? g_tune[5] : (HW > 256 ? 16 * 1024 : 32 * 1024);
  if (B <= 0 || HW <= 0 || C <= 0 || G <= 0 || C % G != 0 || C % kCC != 0) return -1;
  s.B = B; s.HW = HW; s.C = C; s.G = G; s.cpg = C / G;
  if (s.cpg != 10 && s.cpg != 20 && s.cpg != 40 && s.cpg != 80) return -1;
  if (kCC % s.cpg != 0) return -1;
  s.vpc = kVPC; s.chunks = C / kCC;
  const int max_threads = bwd ? (g_tune[1] < 320 ? g_tune[1] : 320) : (g_tune[0] < kMaxThreads ? g_tune[0] : kMaxThreads);
  s.S = 1;
  while (s.S < g_tune[3] && (long long)((HW + s.S - 1) / s.S) * kCC * 2 > slab_target) s.S *= 2;
  s.rp = (HW + s.S - 1) / s.S;
  if ((long long)(s.S - 1) * s.rp >= HW) return -1;          // every rank must own at least one row
  double best = -1.0;
  const int r_unit = kVPC == 20 ? 8 : 16;                    // TU = vpc * r must be a whole number of warps
  for (int r = r_unit; r * kVPC <= max_threads; r += r_unit) {
    const int iters = (s.rp + r - 1) / r;
    if (iters <= kMaxIters) {
      const int tu = kVPC * r;
      int u = s.S > 1 ? 1 : max_threads / tu;
      if (u > B * s.chunks) u = B * s.chunks;
      const double eff = (double)s.rp / ((double)iters * r);
      const double score = eff * 1000.0 + (double)(tu * u) / max_threads * 10.0 + iters * 0.1;
      if (score > best) {
        best = score;
        s.R = r; s.TU = tu; s.U = u; s.iters = iters; s.threads = tu * u;
      }
    }
    if (r >= s.rp) break;
  }
  s.pf = 0;
  return best < 0.0 ? -1 : 0;
}

static size_t fwd_smem(const Shape& s);
static size_t bwd_smem(const Shape& s);
// The CTAs of a grid start in blockIdx order, so the CTA that takes over a slot of this SM generation is (about) one
// residency ahead: resident CTAs = 148 SMs x min(shared-memory, register, warp limits).  Each CTA asks L2 for that
// CTA's rows right after it has requested its own, so the later CTA's load phase is an L2 hit and DRAM keeps
// streaming while the resident CTAs reduce and store (the kernels are latency bound, profiles/r01_ncu_groupnorm_v3.md).
static void plan_prefetch(Shape& s, int bwd) {
  if (!g_tune_env_read) {
    g_tune_env_read = true;
    if (const char* e = getenv("PSG_GN_PREFETCH")) g_tune[6] = atoi(e);
  }
  s.pf = 0;
  static const int strict = getenv("PSG_GN_STRICT_BARRIER") ? atoi(getenv("PSG_GN_STRICT_BARRIER")) : 0;
  s.strict = strict;
  // accumulate variants (43 of a step's 61 backward norms) re-read dx in the store pass, a dependent global load per row with
  // nothing to hide it; requesting those lines at kernel start makes them L2 hits (PSG_GN_PREFETCH_DX=0 for A/B)
  static const int pf_dx = getenv("PSG_GN_PREFETCH_DX") ? atoi(getenv("PSG_GN_PREFETCH_DX")) : 1;
  s.pf_dx = bwd ? pf_dx : 0;
  if (g_tune[6] <= 0) return;
  // forward only by default: the backward is not waiting for DRAM (same time with and without, profiles/
  // r02_bench_groupnorm_prefetch_sweep.txt) and the requests cost it 7 % more instructions; PSG_GN_PREFETCH_BWD=1 for A/B
  static const int bwd_too = getenv("PSG_GN_PREFETCH_BWD") ? atoi(getenv("PSG_GN_PREFETCH_BWD")) : 0;
  if (bwd && !bwd_too) return;
  int resident = g_tune[6];
  if (resident == 1) {
    const size_t smem = (bwd ? bwd_smem(s) : fwd_smem(s)) + 1024;
    int per_sm = (int)((size_t)227 * 1024 / smem);
    const int by_regs = 65536 / ((bwd ? 96 : 64) * ((s.threads + 31) / 32 * 32));
    if (per_sm > by_regs) per_sm = by_regs;
    if (per_sm > 2048 / s.threads) per_sm = 2048 / s.threads;
    if (per_sm < 1) per_sm = 1;
    resident = 148 * per_sm;
  }
  const int per_cta_units = s.S > 1 ? 1 : s.U;
  int pf = resident / (s.S * per_cta_units);           // units one residency ahead
  if (pf < 1) pf = 1;
  s.pf = pf * per_cta_units;                           // whole CTAs ahead (U units per CTA when S == 1)
}

static size_t fwd_smem(const Shape& s) {
  const int kCC = s.vpc * 8, kNP = kCC / 2, ng = kCC / s.cpg;
  return (size_t)s.iters * s.threads * 16 + ((size_t)s.threads * 8 + (size_t)s.U * 2 * kNP + (size_t)s.U * ng * 4) * sizeof(float);
}
static size_t bwd_smem(const Shape& s) {
  const int kCC = s.vpc * 8, ng = kCC / s.cpg;
  return (size_t)2 * s.iters * s.threads * 16 + ((size_t)s.threads * 24 + (size_t)s.U * kCC * 10 + (size_t)s.U * ng * 2) * sizeof(float);
}

// ---- packed fp32 pairs (FFMA2 / FADD2 / FMUL2) ----------------------------------------------------------------------
typedef unsigned long long u64;
__device__ __forceinline__ u64 as_u64(float2 v) { return *reinterpret_cast<u64*>(&v); }
__device__ __forceinline__ float2 as_f2(u64 v) { return *reinterpret_cast<float2*>(&v); }
__device__ __forceinline__ float2 fma2(float2 a, float2 b, float2 c) {
  u64 d;
  asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(as_u64(a)), "l"(as_u64(b)), "l"(as_u64(c)));
  return as_f2(d);
}
__device__ __forceinline__ float2 add2(float2 a, float2 b) {
  u64 d;
  asm("add.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(as_u64(a)), "l"(as_u64(b)));
  return as_f2(d);
}
__device__ __forceinline__ float2 mul2(float2 a, float2 b) {
  u64 d;
  asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(as_u64(a)), "l"(as_u64(b)));
  return as_f2(d);
}
__device__ __forceinline__ void unpack8(const uint4& r, float2 (&v)[4]) {
  const uint32_t w[4] = {r.x, r.y, r.z, r.w};
#pragma unroll
  for (int i = 0; i < 4; ++i) v[i] = make_float2(__uint_as_float(w[i] << 16), __uint_as_float(w[i] & 0xffff0000u));
}
__device__ __forceinline__ uint4 pack8(const float2 (&v)[4]) {
  uint4 r;
  __nv_bfloat162* h = reinterpret_cast<__nv_bfloat162*>(&r);
#pragma unroll
  for (int i = 0; i < 4; ++i) h[i] = __floats2bfloat162_rn(v[i].x, v[i].y);
  return r;
}
__device__ __forceinline__ void cp_async16(uint32_t smem_dst, const void* gmem_src, int src_bytes) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(smem_dst), "l"(gmem_src), "r"(src_bytes) : "memory");
}
__device__ __forceinline__ void cp_async_wait_all() {
  asm volatile("cp.async.commit_group;\n\tcp.async.wait_group 0;" ::: "memory");
}
__device__ __forceinline__ void prefetch_l2(const void* p) { asm volatile("prefetch.global.L2 [%0];" ::"l"(p)); }
__device__ __forceinline__ float tanh_approx(float x) {
  float t;
  asm("tanh.approx.f32 %0, %1;" : "=f"(t) : "f"(x));
  return t;
}
// Cluster barrier, arrive side.  `barrier.cluster.arrive.release` costs every thread a MEMBAR.ALL.GPU (SASS; 10 % of a
// CTA's lifetime in the ncu stall samples, profiles/r01_ncu_groupnorm_v3.md).  The data handed over is the CTA's
// shared-memory partial sums, already ordered before thread 0 by the __syncthreads() that precedes every call, so ONE
// cluster-scope fence by thread 0 followed by relaxed arrivals publishes them (release cumulativity through the CTA
// barrier).  The second barrier of a kernel only says "I have finished reading your shared memory": no fence at all.
__device__ __forceinline__ void cluster_arrive_publish(bool leader, int strict) {
  if (strict) { asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory"); return; }
  if (leader) asm volatile("fence.acq_rel.cluster;" ::: "memory");
  asm volatile("barrier.cluster.arrive.relaxed.aligned;" ::: "memory");
}
__device__ __forceinline__ void cluster_arrive_done(int strict) {
  if (strict) { asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory"); return; }
  asm volatile("barrier.cluster.arrive.relaxed.aligned;" ::: "memory");
}
__device__ __forceinline__ void cluster_wait() { asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory"); }
__device__ __forceinline__ void load8(const float* p, float2 (&v)[4]) {
  const float4 a = __ldg(reinterpret_cast<const float4*>(p)), b = __ldg(reinterpret_cast<const float4*>(p + 4));
  v[0] = make_float2(a.x, a.y); v[1] = make_float2(a.z, a.w); v[2] = make_float2(b.x, b.y); v[3] = make_float2(b.z, b.w);
}

// ---------------------------------------------------------------------------------------------------------------------
// forward:  y = act((x - mean) * rstd * gamma + beta)
// ---------------------------------------------------------------------------------------------------------------------
template <int kVPC, int CPG, bool ACT>
__global__ void __launch_bounds__(kMaxThreads, 2) gn_cluster_fwd_kernel(const __nv_bfloat16* __restrict__ x, long long ld,
                                                                       __nv_bfloat16* __restrict__ y, long long ldy,
                                                                       const float* __restrict__ gamma,
                                                                       const float* __restrict__ beta, Shape s,
                                                                       float* __restrict__ stats, float eps) {
  constexpr int kCC = kVPC * 8, kNP = kCC / 2, NG = kCC / CPG, PPG = CPG / 2;
  extern __shared__ uint4 slab[];                                             // [iters][threads] staged x vectors
  const int T = s.threads, tid = threadIdx.x;
  float* red = reinterpret_cast<float*>(slab + (size_t)s.iters * T);          // [2][U*R][kNP] per-thread pair sums / squares
  float* col = red + (size_t)T * 8;                                           // [U][2][kNP] summed over the rows in flight
  float* part = col + (size_t)s.U * 2 * kNP;                                  // [U][NG][2] group sums (cluster exchange)
  float* st = part + (size_t)s.U * NG * 2;                                    // [U][NG][2] mean, rstd
  int u = 0, tu = tid;
  if (s.U > 1) { u = tid / s.TU; tu = tid - u * s.TU; }
  const int row = tu / kVPC, v = tu - row * kVPC;
  const int rank = s.S > 1 ? (int)(blockIdx.x % s.S) : 0;
  const int unit = s.S > 1 ? (int)(blockIdx.x / s.S) : (int)blockIdx.x * s.U + u;
  const bool active = unit < s.B * s.chunks;
  const int b = active ? unit / s.chunks : 0;
  const int ch = active ? unit - b * s.chunks : 0;
  const int c0 = ch * kCC + v * 8;
  const int row0 = rank * s.rp;
  const int nrows = active ? min(s.rp, s.HW - row0) : 0;
  const uint4* mine = slab + tid;

  {  // the whole slab is requested at once; each thread later reads back only the slots it copied itself
    const __nv_bfloat16* src = x + ((long long)b * s.HW + row0 + row) * ld + c0;
    const long long step = (long long)s.R * ld;
    uint32_t dst = (uint32_t)__cvta_generic_to_shared(slab + tid);
    for (int k = 0, p = row; k < s.iters; ++k, p += s.R, src += step, dst += (uint32_t)T * 16) {
      const bool ok = p < nrows;
      cp_async16(dst, ok ? src : x, ok ? 16 : 0);
    }
    // the same rows of the unit that runs one residency later: one request per 128-byte line of a row piece
    const int unit_pf = unit + s.pf;
    if (s.pf > 0 && (v & 7) == 0 && unit_pf < s.B * s.chunks) {
      const int b2 = unit_pf / s.chunks, ch2 = unit_pf - b2 * s.chunks;
      const __nv_bfloat16* px = x + ((long long)b2 * s.HW + row0 + row) * ld + ch2 * kCC + v * 8;
      for (int p = row; p < nrows; p += s.R, px += step) prefetch_l2(px);
    }
  }
  float2 ga[4], be[4];      // affine parameters: fetched while the slab is in flight
  load8(gamma + c0, ga);
  load8(beta + c0, be);
  cp_async_wait_all();
  {
    float2 su[4], sq[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) { su[i] = make_float2(0.f, 0.f); sq[i] = make_float2(0.f, 0.f); }
#pragma unroll 2
    for (int k = 0; k < s.iters; ++k) {
      float2 f[4];
      unpack8(mine[(size_t)k * T], f);
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        su[i] = add2(su[i], f[i]);
        sq[i] = fma2(f[i], f[i], sq[i]);
      }
    }
    float* r0 = red + (size_t)(u * s.R + row) * kNP + v * 4;
    *reinterpret_cast<float4*>(r0) = make_float4(su[0].x + su[0].y, su[1].x + su[1].y, su[2].x + su[2].y, su[3].x + su[3].y);
    *reinterpret_cast<float4*>(r0 + (size_t)T * 4) = make_float4(sq[0].x + sq[0].y, sq[1].x + sq[1].y, sq[2].x + sq[2].y, sq[3].x + sq[3].y);
  }
  __syncthreads();
  for (int i = tu; i < 2 * kNP; i += s.TU) {           // fold the R rows in flight, fixed order
    const int stat = i / kNP, p = i - stat * kNP;
    const float* src = red + (size_t)stat * T * 4 + (size_t)u * s.R * kNP + p;
    float t0 = 0.f, t1 = 0.f, t2 = 0.f, t3 = 0.f;
    for (int r = 0; r < s.R; r += 4) {                 // R is a multiple of 8
      t0 += src[(r + 0) * kNP];
      t1 += src[(r + 1) * kNP];
      t2 += src[(r + 2) * kNP];
      t3 += src[(r + 3) * kNP];
    }
    col[(u * 2 + stat) * kNP + p] = (t0 + t1) + (t2 + t3);
  }
  __syncthreads();
  if (tu < 2 * NG) {                                   // a channel pair never straddles a group (CPG is even)
    const int g = tu >> 1, stat = tu & 1;
    const float* src = col + (u * 2 + stat) * kNP + g * PPG;
    float t = 0.f;
#pragma unroll
    for (int k = 0; k < PPG; ++k) t += src[k];
    part[(u * NG + g) * 2 + stat] = t;
  }
  __syncthreads();
  const float inv_m = 1.f / ((float)CPG * (float)s.HW);
  if (s.S > 1) {
    cluster_arrive_publish(tid == 0, s.strict);
    cluster_wait();                                    // every CTA's part[] is visible cluster-wide
    if (tid < NG) {
      cg::cluster_group cluster = cg::this_cluster();
      float su = 0.f, sq = 0.f;
      for (int r = 0; r < s.S; ++r) {                  // rank order: every CTA of the cluster gets identical bits
        const float* rp = cluster.map_shared_rank(part, r);
        su += rp[tid * 2];
        sq += rp[tid * 2 + 1];
      }
      const float mean = su * inv_m;
      const float rstd = rsqrtf(fmaxf(sq * inv_m - mean * mean, 0.f) + eps);
      st[tid * 2] = mean;
      st[tid * 2 + 1] = rstd;
      if (rank == 0) *reinterpret_cast<float2*>(stats + ((long long)b * s.G + ch * NG + tid) * 2) = make_float2(mean, rstd);
    }
  } else if (tu < NG) {
    const float mean = part[(u * NG + tu) * 2] * inv_m;
    const float rstd = rsqrtf(fmaxf(part[(u * NG + tu) * 2 + 1] * inv_m - mean * mean, 0.f) + eps);
    st[(u * NG + tu) * 2] = mean;
    st[(u * NG + tu) * 2 + 1] = rstd;
    if (active) *reinterpret_cast<float2*>(stats + ((long long)b * s.G + ch * NG + tu) * 2) = make_float2(mean, rstd);
  }
  __syncthreads();
  if (s.S > 1) cluster_arrive_done(s.strict);                  // done reading the peers' shared memory (waited on before exit)
  if (active) {
    float2 sc[4], sh[4];      // n = x * sc + sh;  with SiLU: h = n / 2 and y = h + h * tanh(h)
    {
      constexpr float half = ACT ? 0.5f : 1.f;
      float a[8], c[8];
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        const int gi = (v * 8 + j) / CPG;
        const float2 m = *reinterpret_cast<const float2*>(st + (u * NG + gi) * 2);
        const float gj = (j & 1) ? ga[j >> 1].y : ga[j >> 1].x, bj = (j & 1) ? be[j >> 1].y : be[j >> 1].x;
        a[j] = m.y * gj * half;
        c[j] = bj * half - m.x * a[j];
      }
#pragma unroll
      for (int i = 0; i < 4; ++i) { sc[i] = make_float2(a[2 * i], a[2 * i + 1]); sh[i] = make_float2(c[2 * i], c[2 * i + 1]); }
    }
    __nv_bfloat16* dst = y + ((long long)b * s.HW + row0 + row) * ldy + c0;
    const long long step = (long long)s.R * ldy;
#pragma unroll 2
    for (int k = 0, p = row; k < s.iters; ++k, p += s.R, dst += step) {
      if (p >= nrows) break;
      float2 f[4];
      unpack8(mine[(size_t)k * T], f);
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const float2 h = fma2(f[i], sc[i], sh[i]);
        f[i] = ACT ? fma2(h, make_float2(tanh_approx(h.x), tanh_approx(h.y)), h) : h;
      }
      *reinterpret_cast<uint4*>(dst) = pack8(f);
    }
  }
  __syncwarp();
  if (s.S > 1) cluster_wait();
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
// The pixel loop accumulates the raw moments  s1, m2 = sum dn*x, m3 = sum x  (one packed op each) and the per-channel
// epilogue converts them:  s2 = rs*m2 - mean*rs*s1,  s3 = rs*m3 - HW*mean*rs.
// SiLU'(n) with h = n/2, t = tanh(h), sig = (1 + t)/2:  sig * (1 + n*(1 - sig)) = sig + sig * h * (1 - t)
// Per-channel constants are computed once per CTA by the first CC threads and read back as vectors.
// ---------------------------------------------------------------------------------------------------------------------
__device__ __forceinline__ void lds8(const float* p, float2 (&v)[4]) {
  const float4 a = *reinterpret_cast<const float4*>(p), b = *reinterpret_cast<const float4*>(p + 4);
  v[0] = make_float2(a.x, a.y); v[1] = make_float2(a.z, a.w); v[2] = make_float2(b.x, b.y); v[3] = make_float2(b.z, b.w);
}

template <int kVPC, int CPG, bool ACT, bool ACCUM>
__global__ void __launch_bounds__(kMaxThreads / 2 + 64, 2) gn_cluster_bwd_kernel(
    const __nv_bfloat16* __restrict__ dy, long long lddy, const __nv_bfloat16* __restrict__ x, long long ld,
    __nv_bfloat16* __restrict__ dx, long long lddx, const float* __restrict__ gamma, const float* __restrict__ beta,
    const float* __restrict__ stats, Shape s, float* __restrict__ partial, float* __restrict__ colsum_out, long long ld_colsum) {
  constexpr int kCC = kVPC * 8, NG = kCC / CPG;
  extern __shared__ uint4 slab[];                       // [2][iters][threads] staged x and dy vectors
  const int T = s.threads, tid = threadIdx.x;
  const uint4* slab_x = slab + tid;
  uint4* slab_d = slab + (size_t)s.iters * T + tid;
  float* acc = reinterpret_cast<float*>(slab + (size_t)2 * s.iters * T);   // [3][U*R][kCC] per-thread channel sums
  float* chan = acc + (size_t)T * 24;                                       // [U][3][kCC] this CTA's per-channel sums
  float* tot = chan + (size_t)s.U * kCC * 3;                                // [U][3][kCC] sums over the cluster
  float* cst = tot + (size_t)s.U * kCC * 3;                                 // [U][4][kCC] per-channel constants
  float* grp = cst + (size_t)s.U * kCC * 4;                                 // [U][NG][2]
  int u = 0, tu = tid;
  if (s.U > 1) { u = tid / s.TU; tu = tid - u * s.TU; }
  const int row = tu / kVPC, v = tu - row * kVPC;
  const int rank = s.S > 1 ? (int)(blockIdx.x % s.S) : 0;
  const int unit = s.S > 1 ? (int)(blockIdx.x / s.S) : (int)blockIdx.x * s.U + u;
  const bool active = unit < s.B * s.chunks;
  const int b = active ? unit / s.chunks : 0;
  const int ch = active ? unit - b * s.chunks : 0;
  const int c0 = ch * kCC + v * 8;
  const int row0 = rank * s.rp;
  const int nrows = active ? min(s.rp, s.HW - row0) : 0;
  float* my_cst = cst + (size_t)u * 4 * kCC;

  {
    const __nv_bfloat16* sx = x + ((long long)b * s.HW + row0 + row) * ld + c0;
    const __nv_bfloat16* sd = dy + ((long long)b * s.HW + row0 + row) * lddy + c0;
    const long long stepx = (long long)s.R * ld, stepd = (long long)s.R * lddy;
    uint32_t dst = (uint32_t)__cvta_generic_to_shared(slab + tid);
    const uint32_t off_d = (uint32_t)s.iters * T * 16;
    for (int k = 0, p = row; k < s.iters; ++k, p += s.R, sx += stepx, sd += stepd, dst += (uint32_t)T * 16) {
      const bool ok = p < nrows;
      cp_async16(dst, ok ? sx : x, ok ? 16 : 0);
      cp_async16(dst + off_d, ok ? sd : dy, ok ? 16 : 0);
    }
    if (ACCUM && s.pf_dx && (v & 7) == 0) {      // this CTA's own dx rows (read back in the store pass): one request per 128-byte line
      const __nv_bfloat16* po = dx + ((long long)b * s.HW + row0 + row) * lddx + c0;
      const long long stepo = (long long)s.R * lddx;
      for (int p = row; p < nrows; p += s.R, po += stepo) prefetch_l2(po);
    }
    // the same rows of the unit that runs one residency later: one request per 128-byte line of a row piece
    const int unit_pf = unit + s.pf;
    if (s.pf > 0 && (v & 7) == 0 && unit_pf < s.B * s.chunks) {
      const int b2 = unit_pf / s.chunks, ch2 = unit_pf - b2 * s.chunks;
      const __nv_bfloat16* px = x + ((long long)b2 * s.HW + row0 + row) * ld + ch2 * kCC + v * 8;
      const __nv_bfloat16* pd = dy + ((long long)b2 * s.HW + row0 + row) * lddy + ch2 * kCC + v * 8;
      for (int p = row; p < nrows; p += s.R, px += stepx, pd += stepd) {
        prefetch_l2(px);
        prefetch_l2(pd);
      }
    }
  }
  // per-channel constants while the slab is in flight:  rs, nmr = -mean*rs,  h = x*ah + bh (= n/2)
  float gam = 0.f;
  if (tu < kCC) {
    const int c = ch * kCC + tu;
    const float2 m = __ldg(reinterpret_cast<const float2*>(stats + ((long long)b * s.G + c / CPG) * 2));
    gam = __ldg(gamma + c);
    const float a = m.y * gam * 0.5f;
    my_cst[tu] = m.y;
    my_cst[kCC + tu] = -m.x * m.y;
    my_cst[2 * kCC + tu] = a;
    my_cst[3 * kCC + tu] = __ldg(beta + c) * 0.5f - m.x * a;
  }
  cp_async_wait_all();
  __syncthreads();
  {
    float2 ah[4], bh[4];
    if (ACT) {
      lds8(my_cst + 2 * kCC + v * 8, ah);
      lds8(my_cst + 3 * kCC + v * 8, bh);
    }
    float2 a1[4], a2[4], a3[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) { a1[i] = make_float2(0.f, 0.f); a2[i] = a1[i]; a3[i] = a1[i]; }
    const float2 half2 = make_float2(0.5f, 0.5f), one2 = make_float2(1.f, 1.f), mone2 = make_float2(-1.f, -1.f);
#pragma unroll 2
    for (int k = 0; k < s.iters; ++k) {
      float2 fx[4], fd[4];
      unpack8(slab_x[(size_t)k * T], fx);
      unpack8(slab_d[(size_t)k * T], fd);
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        float2 dn = fd[i];                                // x == dy == 0 on padded slots
        if (ACT) {
          const float2 h = fma2(fx[i], ah[i], bh[i]);
          const float2 t = make_float2(tanh_approx(h.x), tanh_approx(h.y));
          const float2 sig = fma2(t, half2, half2);
          const float2 hw = mul2(h, fma2(t, mone2, one2));
          dn = mul2(dn, fma2(sig, hw, sig));
          fd[i] = dn;
        }
        a1[i] = add2(a1[i], dn);
        a2[i] = fma2(dn, fx[i], a2[i]);
        a3[i] = add2(a3[i], fx[i]);
      }
      if (ACT) slab_d[(size_t)k * T] = pack8(fd);     // keep dn (bf16) for the second pass instead of re-evaluating SiLU'
    }
    float* r0 = acc + (size_t)(u * s.R + row) * kCC + v * 8;
    const size_t plane = (size_t)T * 8;
    *reinterpret_cast<float4*>(r0) = make_float4(a1[0].x, a1[0].y, a1[1].x, a1[1].y);
    *reinterpret_cast<float4*>(r0 + 4) = make_float4(a1[2].x, a1[2].y, a1[3].x, a1[3].y);
    *reinterpret_cast<float4*>(r0 + plane) = make_float4(a2[0].x, a2[0].y, a2[1].x, a2[1].y);
    *reinterpret_cast<float4*>(r0 + plane + 4) = make_float4(a2[2].x, a2[2].y, a2[3].x, a2[3].y);
    *reinterpret_cast<float4*>(r0 + 2 * plane) = make_float4(a3[0].x, a3[0].y, a3[1].x, a3[1].y);
    *reinterpret_cast<float4*>(r0 + 2 * plane + 4) = make_float4(a3[2].x, a3[2].y, a3[3].x, a3[3].y);
  }
  __syncthreads();
  // per-channel fold over the R rows in flight, fixed order: thread (stat, c) of the unit
  for (int idx = tu; idx < 3 * kCC; idx += s.TU) {
    const int stat = idx / kCC, c = idx - stat * kCC;
    const float* src = acc + (size_t)stat * T * 8 + (size_t)u * s.R * kCC + c;
    float t0 = 0.f, t1 = 0.f, t2 = 0.f, t3 = 0.f;
    for (int r = 0; r < s.R; r += 4) {                 // R is a multiple of 8
      t0 += src[(r + 0) * kCC];
      t1 += src[(r + 1) * kCC];
      t2 += src[(r + 2) * kCC];
      t3 += src[(r + 3) * kCC];
    }
    chan[(u * 3 + stat) * kCC + c] = (t0 + t1) + (t2 + t3);
  }
  __syncthreads();
  // totals over the cluster (rank order: identical bits in every CTA), converted to the centred sums s1, s2, s3
  float s1 = 0.f, s2 = 0.f, s3 = 0.f;
  if (s.S > 1) {
    cluster_arrive_publish(tid == 0, s.strict);
    cluster_wait();                                       // every CTA's chan[] is visible cluster-wide
    if (tid < kCC) {
      cg::cluster_group cluster = cg::this_cluster();
      float m2 = 0.f, m3 = 0.f;
      for (int r = 0; r < s.S; ++r) {
        const float* rc = cluster.map_shared_rank(chan, r);
        s1 += rc[tid];
        m2 += rc[kCC + tid];
        m3 += rc[2 * kCC + tid];
      }
      s2 = m2;
      s3 = m3;
    }
  } else if (tu < kCC) {
    s1 = chan[(u * 3 + 0) * kCC + tu];
    s2 = chan[(u * 3 + 1) * kCC + tu];
    s3 = chan[(u * 3 + 2) * kCC + tu];
  }
  float rs_c = 0.f, nmr_c = 0.f;
  if (tu < kCC) {
    rs_c = my_cst[tu];
    nmr_c = my_cst[kCC + tu];
    s2 = fmaf(rs_c, s2, nmr_c * s1);
    s3 = fmaf(rs_c, s3, (float)s.HW * nmr_c);
    tot[(u * 3 + 0) * kCC + tu] = gam * s1;
    tot[(u * 3 + 1) * kCC + tu] = gam * s2;
  }
  __syncthreads();
  if (s.S > 1) cluster_arrive_done(s.strict);                     // done reading the peers' shared memory (waited on before exit)
  if (tu < 2 * NG) {
    const int stat = tu & 1, g = tu >> 1;
    const float* sv = tot + (u * 3 + stat) * kCC + g * CPG;
    float t = 0.f;
#pragma unroll 10
    for (int c = 0; c < CPG; ++c) t += sv[c];
    grp[(u * NG + g) * 2 + stat] = t;
  }
  __syncthreads();
  // dx = dn*k1 - k2 - xh*k3 with k1 = rstd*gamma, k2 = rstd*A/m, k3 = rstd*Bs/m;  xh = x*rs + nmr  =>
  // dx = dn*k1 + x*kx + k0  with  kx = -rs*k3,  k0 = -(nmr*k3 + k2)
  if (tu < kCC) {
    const float inv_m = 1.f / ((float)CPG * (float)s.HW);
    const float2 ab = *reinterpret_cast<const float2*>(grp + (u * NG + tu / CPG) * 2);
    const float k2 = rs_c * ab.x * inv_m, k3 = rs_c * ab.y * inv_m;
    my_cst[tu] = rs_c * gam;
    my_cst[kCC + tu] = -rs_c * k3;
    my_cst[2 * kCC + tu] = -fmaf(nmr_c, k3, k2);
    if (active && rank == 0) {
      const int c = ch * kCC + tu;
      const float cs = rs_c * (gam * s1 - ((float)s.HW * ab.x + s3 * ab.y) * inv_m);
      float* o = partial + ((long long)b * s.C + c) * 3;
      o[0] = s1;
      o[1] = s2;
      o[2] = cs;
      if (colsum_out) colsum_out[(long long)b * ld_colsum + c] = cs;
    }
  }
  __syncthreads();
  if (active) {
    float2 k1[4], kx[4], k0[4];
    lds8(my_cst + v * 8, k1);
    lds8(my_cst + kCC + v * 8, kx);
    lds8(my_cst + 2 * kCC + v * 8, k0);
    __nv_bfloat16* dst = dx + ((long long)b * s.HW + row0 + row) * lddx + c0;
    const long long step = (long long)s.R * lddx;
#pragma unroll 2
    for (int k = 0, p = row; k < s.iters; ++k, p += s.R, dst += step) {
      if (p >= nrows) break;
      float2 fx[4], fd[4], fo[4];
      unpack8(slab_x[(size_t)k * T], fx);
      unpack8(slab_d[(size_t)k * T], fd);
      if (ACCUM) unpack8(*reinterpret_cast<const uint4*>(dst), fo);
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const float2 d = fma2(fd[i], k1[i], fma2(fx[i], kx[i], k0[i]));      // fd holds dn
        fo[i] = ACCUM ? add2(fo[i], d) : d;
      }
      *reinterpret_cast<uint4*>(dst) = pack8(fo);
    }
  }
  __syncwarp();
  if (s.S > 1) cluster_wait();
}

template <typename Kern>
static int configure(const char* what, Kern kernel) {
  cudaError_t e = cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kSmemLimit);
  if (e == cudaSuccess) e = cudaFuncSetAttribute(kernel, cudaFuncAttributeNonPortableClusterSizeAllowed, 1);
  if (e != cudaSuccess) { psg_set_error("%s: cudaFuncSetAttribute: %s", what, cudaGetErrorString(e)); return PSG_ERR_CUDA; }
  return PSG_OK;
}

template <typename Kern, typename... Args>
static int launch(const char* what, Kern kernel, int grid, int threads, size_t smem, int cluster, cudaStream_t stream, Args... args) {
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3((unsigned)grid);
  cfg.blockDim = dim3((unsigned)threads);
  cfg.dynamicSmemBytes = smem;
  cfg.stream = stream;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = (unsigned)cluster;
  attr[0].val.clusterDim.y = 1;
  attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  cudaError_t e = cudaLaunchKernelEx(&cfg, kernel, args...);
  if (e != cudaSuccess) { psg_set_error("%s: launch failed: %s", what, cudaGetErrorString(e)); return PSG_ERR_CUDA; }
  ++g_psg_launch_count;
  return PSG_OK;
}

template <int kVPC, int CPG, bool ACT>
static int run_fwd(const Shape& s, const void* x, long long ld_x, void* y, long long ld_y, const float* gamma, const float* beta,
                   float* stats, float eps, cudaStream_t stream) {
  const int units = s.B * s.chunks;
  const int grid = s.S > 1 ? units * s.S : (units + s.U - 1) / s.U;
  static bool configured = false;          // one flag per kernel instantiation
  if (!configured) {
    const int rc = configure("psg_groupnorm_fused_fwd(cluster)", gn_cluster_fwd_kernel<kVPC, CPG, ACT>);
    if (rc != PSG_OK) return rc;
    configured = true;
  }
  return launch("psg_groupnorm_fused_fwd(cluster)", gn_cluster_fwd_kernel<kVPC, CPG, ACT>, grid, s.threads, fwd_smem(s), s.S, stream,
                (const __nv_bfloat16*)x, ld_x, (__nv_bfloat16*)y, ld_y, gamma, beta, s, stats, eps);
}
template <int kVPC, int CPG, bool ACT, bool ACCUM>
static int run_bwd2(const Shape& s, const void* dy, long long ld_dy, const void* x, long long ld_x, void* dx, long long ld_dx,
                    const float* gamma, const float* beta, const float* stats, float* partial, float* dx_colsum, long long ld_colsum,
                    cudaStream_t stream) {
  const int units = s.B * s.chunks;
  const int grid = s.S > 1 ? units * s.S : (units + s.U - 1) / s.U;
  static bool configured = false;          // one flag per kernel instantiation
  if (!configured) {
    const int rc = configure("psg_groupnorm_fused_bwd(cluster)", gn_cluster_bwd_kernel<kVPC, CPG, ACT, ACCUM>);
    if (rc != PSG_OK) return rc;
    configured = true;
  }
  return launch("psg_groupnorm_fused_bwd(cluster)", gn_cluster_bwd_kernel<kVPC, CPG, ACT, ACCUM>, grid, s.threads, bwd_smem(s), s.S, stream,
                (const __nv_bfloat16*)dy, ld_dy, (const __nv_bfloat16*)x, ld_x, (__nv_bfloat16*)dx, ld_dx, gamma, beta, stats, s,
                partial, dx_colsum, ld_colsum);
}
template <int kVPC, int CPG, bool ACT>
static int run_bwd(const Shape& s, const void* dy, long long ld_dy, const void* x, long long ld_x, void* dx, long long ld_dx,
                   const float* gamma, const float* beta, const float* stats, float* partial, float* dx_colsum, long long ld_colsum,
                   int accumulate_dx, cudaStream_t stream) {
  return accumulate_dx ? run_bwd2<kVPC, CPG, ACT, true>(s, dy, ld_dy, x, ld_x, dx, ld_dx, gamma, beta, stats, partial, dx_colsum, ld_colsum, stream)
                       : run_bwd2<kVPC, CPG, ACT, false>(s, dy, ld_dy, x, ld_x, dx, ld_dx, gamma, beta, stats, partial, dx_colsum, ld_colsum, stream);
}

}  // namespace gnc

#define GNC_CASE(V, P, ...)                                                              \
  case (V) * 1000 + (P) * 2: return gnc::FN<V, P, false>(__VA_ARGS__);                    \
  case (V) * 1000 + (P) * 2 + 1: return gnc::FN<V, P, true>(__VA_ARGS__);
#define GNC_DISPATCH(...)                                                                \
  switch (s.vpc * 1000 + s.cpg * 2 + (act ? 1 : 0)) {                                    \
    GNC_CASE(20, 10, __VA_ARGS__)                                                        \
    GNC_CASE(20, 20, __VA_ARGS__)                                                        \
    GNC_CASE(20, 40, __VA_ARGS__)                                                        \
    GNC_CASE(20, 80, __VA_ARGS__)                                                        \
    GNC_CASE(10, 10, __VA_ARGS__)                                                        \
    GNC_CASE(10, 20, __VA_ARGS__)                                                        \
    GNC_CASE(10, 40, __VA_ARGS__)                                                        \
    GNC_CASE(10, 80, __VA_ARGS__)                                                        \
    default: return PSG_ERR_UNSUPPORTED;                                                 \
  }

int gnc_tune(int which, int value) {
  if (which < 0 || which > 6) return -1;
  if (which == 4 && value > 0 && value != 10 && value != 20) return -1;
  const int prev = gnc::g_tune[which];
  if (which == 6 && value >= 0) gnc::g_tune_env_read = true;   // an explicit setting wins over the environment
  if (value > 0 || (value == 0 && which >= 4)) gnc::g_tune[which] = value;
  return prev;
}

int gnc_supported(int B, int HW, int C, int G, int bwd) {
  gnc::Shape s;
  if (gnc::plan(s, B, HW, C, G, bwd) != 0) return 0;
  return (bwd ? gnc::bwd_smem(s) : gnc::fwd_smem(s)) <= gnc::kSmemLimit ? 1 : 0;
}

int gnc_plan(int B, int HW, int C, int G, int bwd, int* out) {
  gnc::Shape s;
  if (gnc::plan(s, B, HW, C, G, bwd) != 0) return -1;
  out[0] = s.vpc * 8; out[1] = s.S; out[2] = s.rp; out[3] = s.R; out[4] = s.TU; out[5] = s.U; out[6] = s.iters;
  out[7] = (int)(bwd ? gnc::bwd_smem(s) : gnc::fwd_smem(s));
  return 0;
}

int gnc_fwd(const void* x, long long ld_x, void* y, long long ld_y, const float* gamma, const float* beta, float* stats, int B,
            int HW, int C, int G, float eps, int act, cudaStream_t stream) {
  gnc::Shape s;
  if (gnc::plan(s, B, HW, C, G, 0) != 0 || gnc::fwd_smem(s) > gnc::kSmemLimit) return PSG_ERR_UNSUPPORTED;
  gnc::plan_prefetch(s, 0);
#define FN run_fwd
  GNC_DISPATCH(s, x, ld_x, y, ld_y, gamma, beta, stats, eps, stream)
#undef FN
}

int gnc_bwd(const void* dy, long long ld_dy, const void* x, long long ld_x, void* dx, long long ld_dx, const float* gamma,
            const float* beta, const float* stats, float* partial, float* dx_colsum, long long ld_colsum, int B, int HW, int C, int G,
            int act, int accumulate_dx, cudaStream_t stream) {
  gnc::Shape s;
  if (gnc::plan(s, B, HW, C, G, 1) != 0 || gnc::bwd_smem(s) > gnc::kSmemLimit) return PSG_ERR_UNSUPPORTED;
  gnc::plan_prefetch(s, 1);
#define FN run_bwd
  GNC_DISPATCH(s, dy, ld_dy, x, ld_x, dx, ld_dx, gamma, beta, stats, partial, dx_colsum, ld_colsum, accumulate_dx, stream)
#undef FN
}
