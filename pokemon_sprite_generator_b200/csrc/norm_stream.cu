// Streaming two-phase GroupNorm(+SiLU) backward for tensors larger than the L2 (bf16, token-major), ONE launch.
//
// Reference ops replaced: the autograd backward of nn.GroupNorm (+ F.silu) at the 27x27 and 14x14 levels
// (src/models/unet.py:79,89,115,127,156-157,214,231,397-398).  SURVEY.md section 2.1 K5; bound by HBM (3 N elements).
//
// Why a third GroupNorm backward.  The single-pass cluster kernel (norm_cluster.cu) keeps a unit's x and dy in shared
// memory between its two passes: 49 KB and 96 registers allow 4 CTAs x 5 warps per SM, and ncu shows it latency bound
// (24 % of the stall samples on barriers, 16 % on shared-memory loads, 2.9 TB/s = 44 % of the measured copy peak at
// 729 x 320, whether or not its loads hit L2: profiles/r02_ncu_groupnorm_cluster_bwd.md).  Here nothing is staged:
//   phase 0 ("stats")  CTA (sample b, pixel chunk) streams x and dy once from DRAM and accumulates, per channel, the raw
//                      moments  s1 = sum dn,  m2 = sum dn*x,  m3 = sum x  (dn = dy * SiLU'(n)) in registers; the row lanes
//                      of the CTA are folded in fixed order and written to part[b][chunk][3][C]; then ready[b] += 1.
//   phase 1 ("apply")  CTA (b, chunk) waits for ready[b] == chunks, folds the chunks' moments (fixed order), forms the group
//                      sums and the per-channel constants of  dx = dn*k1 + x*kx + k0  in shared memory, and streams x
//                      and dy a second time -- out of L2 -- writing dx.
// The grid is ordered  [stats of sample group 0][apply of group 0][stats of group 1][apply of group 1] ...  with a group
// sized so that its x and dy (48 MB by default) stay L2-resident between the two phases: DRAM traffic stays at the
// algorithmic 3 N.  CTAs start in blockIdx order, so every CTA an apply CTA waits for is already resident or done (the
// wait is bounded anyway: a protocol fault raises a flag, it never hangs the GPU).  Both phases are plain streaming loops
// (4 rows in flight per thread, no barrier inside), every reduction is fixed-order (run-to-run bit-identical).
#include "norm_stream.h"
#include "psg_common.cuh"

namespace gns {

constexpr int kThreadsMax = 320;
constexpr size_t kSmemLimit = 112 * 1024;

__device__ int g_timeout_flag = 0;

struct Params {
  const __nv_bfloat16* dy; long long lddy;
  const __nv_bfloat16* x; long long ldx;
  __nv_bfloat16* dx; long long lddx;
  const float* gamma; const float* beta; const float* stats;
  float* part;                 // [B][nchunks][3][C] raw moments per pixel chunk
  int* ready;                  // [B] chunks of the sample whose moments are written (zeroed by the launcher)
  float* partial;              // [B][C][3] = {s1, s2, sum_pix dx} for the parameter-gradient fold
  float* colsum_out; long long ld_colsum;
  int B, HW, C, G, cpg;
  int V, RL, T;                // 16-byte vectors per row, row lanes per CTA, threads = V * RL
  int rows, nchunks, gs;       // pixel rows per chunk, chunks per sample, samples per L2 group
};

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
__device__ __forceinline__ float tanh_approx(float x) {
  float t;
  asm("tanh.approx.f32 %0, %1;" : "=f"(t) : "f"(x));
  return t;
}
__device__ __forceinline__ void lds8(const float* p, float2 (&v)[4]) {
  const float4 a = *reinterpret_cast<const float4*>(p), b = *reinterpret_cast<const float4*>(p + 4);
  v[0] = make_float2(a.x, a.y); v[1] = make_float2(a.z, a.w); v[2] = make_float2(b.x, b.y); v[3] = make_float2(b.z, b.w);
}
// dn = dy * SiLU'(n) with h = n/2 = x*ah + bh, t = tanh(h), sig = (1 + t)/2:  SiLU'(n) = sig + sig * h * (1 - t)
template <bool ACT>
__device__ __forceinline__ void silu_grad(const float2 (&fx)[4], float2 (&fd)[4], const float2 (&ah)[4], const float2 (&bh)[4]) {
  if (!ACT) return;
  const float2 half2 = make_float2(0.5f, 0.5f), one2 = make_float2(1.f, 1.f), mone2 = make_float2(-1.f, -1.f);
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const float2 h = fma2(fx[i], ah[i], bh[i]);
    const float2 t = make_float2(tanh_approx(h.x), tanh_approx(h.y));
    const float2 sig = fma2(t, half2, half2);
    const float2 hw = mul2(h, fma2(t, mone2, one2));
    fd[i] = mul2(fd[i], fma2(sig, hw, sig));
  }
}

template <bool ACT, bool ACCUM>
__global__ void __launch_bounds__(kThreadsMax, 2) gn_stream_bwd_kernel(const Params p) {
  extern __shared__ float smem[];
  const int tid = threadIdx.x;
  const int per_group = p.gs * p.nchunks;
  const int grp_i = (int)blockIdx.x / (2 * per_group);
  int within = (int)blockIdx.x - grp_i * 2 * per_group;
  const int phase = within >= per_group ? 1 : 0;
  if (phase) within -= per_group;
  const int b = grp_i * p.gs + within / p.nchunks;
  const int chunk = within % p.nchunks;
  if (b >= p.B) return;                                  // ragged last group (both phases agree)
  const int C = p.C;
  const int v = tid % p.V, rl = tid / p.V;
  const int c0 = v * 8;
  const int r0 = chunk * p.rows;
  const int r1 = min(p.HW, r0 + p.rows);
  const int n_it = r0 + rl < r1 ? (r1 - r0 - rl + p.RL - 1) / p.RL : 0;
  const long long row_first = (long long)b * p.HW + r0 + rl;
  const uint4* px = reinterpret_cast<const uint4*>(p.x + row_first * p.ldx + c0);
  const uint4* pd = reinterpret_cast<const uint4*>(p.dy + row_first * p.lddy + c0);
  const long long sx = (long long)p.RL * p.ldx / 8, sd = (long long)p.RL * p.lddy / 8;   // pitches are multiples of 8 elements

  if (phase == 0) {
    // ---- stats: raw moments of this chunk ----
    float2 ah[4], bh[4];
    if (ACT) {
      float a[8], c[8];
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        const int ch = c0 + j;
        const float2 m = __ldg(reinterpret_cast<const float2*>(p.stats + ((long long)b * p.G + ch / p.cpg) * 2));
        a[j] = m.y * __ldg(p.gamma + ch) * 0.5f;
        c[j] = __ldg(p.beta + ch) * 0.5f - m.x * a[j];
      }
#pragma unroll
      for (int i = 0; i < 4; ++i) { ah[i] = make_float2(a[2 * i], a[2 * i + 1]); bh[i] = make_float2(c[2 * i], c[2 * i + 1]); }
    }
    float2 a1[4], a2[4], a3[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) { a1[i] = make_float2(0.f, 0.f); a2[i] = a1[i]; a3[i] = a1[i]; }
    int it = 0;
    for (; it + 4 <= n_it; it += 4, px += 4 * sx, pd += 4 * sd) {      // four rows (128 B per thread) in flight
      uint4 vx[4], vd[4];
#pragma unroll
      for (int k = 0; k < 4; ++k) { vx[k] = __ldcg(px + k * sx); vd[k] = __ldcg(pd + k * sd); }
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        float2 fx[4], fd[4];
        unpack8(vx[k], fx);
        unpack8(vd[k], fd);
        silu_grad<ACT>(fx, fd, ah, bh);
#pragma unroll
        for (int i = 0; i < 4; ++i) { a1[i] = add2(a1[i], fd[i]); a2[i] = fma2(fd[i], fx[i], a2[i]); a3[i] = add2(a3[i], fx[i]); }
      }
    }
    for (; it < n_it; ++it, px += sx, pd += sd) {
      float2 fx[4], fd[4];
      unpack8(__ldcg(px), fx);
      unpack8(__ldcg(pd), fd);
      silu_grad<ACT>(fx, fd, ah, bh);
#pragma unroll
      for (int i = 0; i < 4; ++i) { a1[i] = add2(a1[i], fd[i]); a2[i] = fma2(fd[i], fx[i], a2[i]); a3[i] = add2(a3[i], fx[i]); }
    }
    // fold the row lanes in fixed order: acc[stat][rl][C]
    float* acc = smem;
    {
      float* q0 = acc + (size_t)rl * C + c0;
      const size_t plane = (size_t)p.RL * C;
      *reinterpret_cast<float4*>(q0) = make_float4(a1[0].x, a1[0].y, a1[1].x, a1[1].y);
      *reinterpret_cast<float4*>(q0 + 4) = make_float4(a1[2].x, a1[2].y, a1[3].x, a1[3].y);
      *reinterpret_cast<float4*>(q0 + plane) = make_float4(a2[0].x, a2[0].y, a2[1].x, a2[1].y);
      *reinterpret_cast<float4*>(q0 + plane + 4) = make_float4(a2[2].x, a2[2].y, a2[3].x, a2[3].y);
      *reinterpret_cast<float4*>(q0 + 2 * plane) = make_float4(a3[0].x, a3[0].y, a3[1].x, a3[1].y);
      *reinterpret_cast<float4*>(q0 + 2 * plane + 4) = make_float4(a3[2].x, a3[2].y, a3[3].x, a3[3].y);
    }
    __syncthreads();
    float* out = p.part + ((long long)b * p.nchunks + chunk) * 3 * C;
    for (int idx = tid; idx < 3 * C; idx += p.T) {
      const int stat = idx / C, c = idx - stat * C;
      const float* src = acc + (size_t)stat * p.RL * C + c;
      float t = 0.f;
      for (int r = 0; r < p.RL; ++r) t += src[(size_t)r * C];
      __stcg(out + idx, t);
    }
    __threadfence();
    __syncthreads();
    if (tid == 0) atomicAdd(p.ready + b, 1);
    return;
  }

  // ---- apply ----
  float* sm_s1 = smem;                 // [C] each
  float* sm_s2 = sm_s1 + C;
  float* sm_s3 = sm_s2 + C;
  float* sm_g1 = sm_s3 + C;
  float* sm_g2 = sm_g1 + C;
  float* coef = sm_g2 + C;             // [5][C]: k1, kx, k0, ah, bh
  float* pp = coef + 5 * (size_t)C;    // [G][8][2]
  float* grp = pp + p.G * 16;          // [G][2]
  if (tid == 0) {                      // every chunk of this sample has published its moments (bounded wait)
    bool ok = false;
    for (uint32_t i = 0; i < (1u << 22); ++i) {
      int r;
      asm volatile("ld.acquire.gpu.global.s32 %0, [%1];" : "=r"(r) : "l"(p.ready + b) : "memory");
      if (r >= p.nchunks) { ok = true; break; }
      __nanosleep(100);
    }
    if (!ok) atomicExch(&g_timeout_flag, 1);
  }
  __syncthreads();
  const float fHW = (float)p.HW;
  for (int c = tid; c < C; c += p.T) {
    const float* src = p.part + (long long)b * p.nchunks * 3 * C + c;
    float s1 = 0.f, m2 = 0.f, m3 = 0.f;
    for (int ch = 0; ch < p.nchunks; ++ch, src += 3 * C) {
      s1 += __ldcg(src);
      m2 += __ldcg(src + C);
      m3 += __ldcg(src + 2 * C);
    }
    const float2 m = __ldg(reinterpret_cast<const float2*>(p.stats + ((long long)b * p.G + c / p.cpg) * 2));
    const float gam = __ldg(p.gamma + c);
    const float nmr = -m.x * m.y;
    const float s2 = fmaf(m.y, m2, nmr * s1);
    sm_s1[c] = s1;
    sm_s2[c] = s2;
    sm_s3[c] = fmaf(m.y, m3, fHW * nmr);
    sm_g1[c] = gam * s1;
    sm_g2[c] = gam * s2;
  }
  __syncthreads();
  for (int idx = tid; idx < p.G * 16; idx += p.T) {      // group sums, two levels, fixed order
    const int stat = idx & 1, j = (idx >> 1) & 7, g = idx >> 4;
    const float* src = stat ? sm_g2 : sm_g1;
    float t = 0.f;
    for (int c = g * p.cpg + j; c < (g + 1) * p.cpg; c += 8) t += src[c];
    pp[idx] = t;
  }
  __syncthreads();
  for (int idx = tid; idx < p.G * 2; idx += p.T) {
    const int stat = idx & 1, g = idx >> 1;
    float t = 0.f;
#pragma unroll
    for (int j = 0; j < 8; ++j) t += pp[(g << 4) + (j << 1) + stat];
    grp[idx] = t;
  }
  __syncthreads();
  {
    const float inv_m = 1.f / ((float)p.cpg * fHW);
    for (int c = tid; c < C; c += p.T) {
      const int g = c / p.cpg;
      const float A = grp[2 * g], Bs = grp[2 * g + 1];
      const float2 m = __ldg(reinterpret_cast<const float2*>(p.stats + ((long long)b * p.G + g) * 2));
      const float gam = __ldg(p.gamma + c);
      const float rs = m.y, nmr = -m.x * m.y;
      const float k2 = rs * A * inv_m, k3 = rs * Bs * inv_m;
      coef[c] = rs * gam;
      coef[C + c] = -rs * k3;
      coef[2 * C + c] = -fmaf(nmr, k3, k2);
      if (ACT) {
        const float a = rs * gam * 0.5f;
        coef[3 * C + c] = a;
        coef[4 * C + c] = __ldg(p.beta + c) * 0.5f - m.x * a;
      }
      if (chunk == 0) {
        const float s1 = sm_s1[c];
        const float cs = rs * (gam * s1 - (fHW * A + sm_s3[c] * Bs) * inv_m);
        float* o = p.partial + ((long long)b * C + c) * 3;
        o[0] = s1;
        o[1] = sm_s2[c];
        o[2] = cs;
        if (p.colsum_out) p.colsum_out[(long long)b * p.ld_colsum + c] = cs;
      }
    }
  }
  __syncthreads();
  float2 k1[4], kx[4], k0[4], ah[4], bh[4];
  lds8(coef + c0, k1);
  lds8(coef + C + c0, kx);
  lds8(coef + 2 * C + c0, k0);
  if (ACT) {
    lds8(coef + 3 * C + c0, ah);
    lds8(coef + 4 * C + c0, bh);
  }
  uint4* po = reinterpret_cast<uint4*>(p.dx + row_first * p.lddx + c0);
  const long long so = (long long)p.RL * p.lddx / 8;
  int it = 0;
  for (; it + 2 <= n_it; it += 2, px += 2 * sx, pd += 2 * sd, po += 2 * so) {
    uint4 vx[2], vd[2], vo[2];
#pragma unroll
    for (int k = 0; k < 2; ++k) {
      vx[k] = __ldcg(px + k * sx);
      vd[k] = __ldcg(pd + k * sd);
      if (ACCUM) vo[k] = __ldcg(po + k * so);
    }
#pragma unroll
    for (int k = 0; k < 2; ++k) {
      float2 fx[4], fd[4], fo[4];
      unpack8(vx[k], fx);
      unpack8(vd[k], fd);
      if (ACCUM) unpack8(vo[k], fo);
      silu_grad<ACT>(fx, fd, ah, bh);
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const float2 d = fma2(fd[i], k1[i], fma2(fx[i], kx[i], k0[i]));
        fo[i] = ACCUM ? add2(fo[i], d) : d;
      }
      __stcg(po + k * so, pack8(fo));
    }
  }
  for (; it < n_it; ++it, px += sx, pd += sd, po += so) {
    float2 fx[4], fd[4], fo[4];
    unpack8(__ldcg(px), fx);
    unpack8(__ldcg(pd), fd);
    if (ACCUM) unpack8(__ldcg(po), fo);
    silu_grad<ACT>(fx, fd, ah, bh);
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const float2 d = fma2(fd[i], k1[i], fma2(fx[i], kx[i], k0[i]));
      fo[i] = ACCUM ? add2(fo[i], d) : d;
    }
    __stcg(po, pack8(fo));
  }
}

// tunables (gns_tune): {bytes of x + dy per L2 group, target pixel rows per chunk (0: by shape), smallest tensor (bytes of x) the
// auto mode hands to these kernels}
// Entry 2 defaults to "never": measured on B200 the two-phase kernel moves exactly the algorithmic bytes (ncu: 239 MB read =
// x and dy once, the apply phase hits L2) but takes 179-193 us at 729 x 320 against 126 us for the cluster kernel -- 1400
// warp-instructions per warp of which the streaming loops are 600 (per-CTA prologues), 2 CTAs per SM at 96 registers
// (profiles/r02_bench_groupnorm_stream.txt, r02_ncu_groupnorm_stream_bwd.md).  It stays selectable (mode 2) and tested.
static long long g_tune[3] = {48ll << 20, 96, 1ll << 60};

static int plan(Params& p, int B, int HW, int C, int G) {
  if (B <= 0 || HW <= 0 || C <= 0 || G <= 0 || C % G != 0 || C % 8 != 0) return -1;
  p.B = B; p.HW = HW; p.C = C; p.G = G; p.cpg = C / G;
  p.V = C / 8;
  if (p.V > kThreadsMax) return -1;
  p.RL = kThreadsMax / p.V;
  p.T = p.V * p.RL;
  const int target = g_tune[1] > 0 ? (int)g_tune[1] : (p.RL * 8 > 48 ? p.RL * 8 : 48);
  p.nchunks = (HW + target / 2) / target;
  if (p.nchunks < 1) p.nchunks = 1;
  p.rows = (HW + p.nchunks - 1) / p.nchunks;
  p.nchunks = (HW + p.rows - 1) / p.rows;
  const long long per_sample = (long long)HW * C * 4;
  long long gs = g_tune[0] / per_sample;
  if (gs < 1) gs = 1;
  if (gs > B) gs = B;
  p.gs = (int)gs;
  return 0;
}
static size_t smem_bytes(const Params& p) {
  const size_t stats = (size_t)3 * p.RL * p.C * sizeof(float);
  const size_t apply = ((size_t)10 * p.C + (size_t)18 * p.G) * sizeof(float);
  return stats > apply ? stats : apply;
}
// workspace layout (floats): [0, B*C*3) partial | part [B][nchunks][3][C] | ready [B] (ints)
static size_t off_part(const Params& p) { return ((size_t)p.B * p.C * 3 + 63) / 64 * 64; }
static size_t off_ready(const Params& p) { return off_part(p) + ((size_t)p.B * p.nchunks * 3 * p.C + 63) / 64 * 64; }

template <bool ACT, bool ACCUM>
static int run(const Params& p, cudaStream_t stream) {
  static bool configured = false;          // one flag per kernel instantiation
  if (!configured) {
    cudaError_t e = cudaFuncSetAttribute(gn_stream_bwd_kernel<ACT, ACCUM>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kSmemLimit);
    if (e != cudaSuccess) { psg_set_error("psg_groupnorm_fused_bwd(stream): cudaFuncSetAttribute: %s", cudaGetErrorString(e)); return PSG_ERR_CUDA; }
    configured = true;
  }
  cudaError_t e = cudaMemsetAsync(p.ready, 0, (size_t)p.B * sizeof(int), stream);
  if (e != cudaSuccess) { psg_set_error("psg_groupnorm_fused_bwd(stream): cudaMemsetAsync: %s", cudaGetErrorString(e)); return PSG_ERR_CUDA; }
  const int groups = (p.B + p.gs - 1) / p.gs;
  const long long grid = (long long)groups * 2 * p.gs * p.nchunks;
  gn_stream_bwd_kernel<ACT, ACCUM><<<(unsigned)grid, p.T, smem_bytes(p), stream>>>(p);
  PSG_CHECK_LAUNCH("psg_groupnorm_fused_bwd(stream)");
  return PSG_OK;
}

}  // namespace gns

long long gns_tune(int which, long long value) {
  if (which < 0 || which > 2) return -1;
  const long long prev = gns::g_tune[which];
  if (value >= 0) gns::g_tune[which] = value;
  return prev;
}

long long gns_workspace_floats(int B, int HW, int C, int G) {
  gns::Params p;
  if (gns::plan(p, B, HW, C, G) != 0 || gns::smem_bytes(p) > gns::kSmemLimit) return 0;
  return (long long)(gns::off_ready(p) + (size_t)B + 64);
}

int gns_wants(int B, int HW, int C) { return (long long)B * HW * C * 2 >= gns::g_tune[2] ? 1 : 0; }

int gns_plan(int B, int HW, int C, int G, int* out) {
  gns::Params p;
  if (gns::plan(p, B, HW, C, G) != 0) return -1;
  out[0] = p.V; out[1] = p.RL; out[2] = p.T; out[3] = p.rows; out[4] = p.nchunks; out[5] = p.gs;
  out[6] = (int)gns::smem_bytes(p); out[7] = (int)(((long long)(B + p.gs - 1) / p.gs) * 2 * p.gs * p.nchunks);
  return 0;
}

int gns_bwd(const void* dy, long long ld_dy, const void* x, long long ld_x, void* dx, long long ld_dx, const float* gamma,
            const float* beta, const float* stats, float* workspace, long long workspace_floats, float* dx_colsum, long long ld_colsum,
            int B, int HW, int C, int G, int act, int accumulate_dx, cudaStream_t stream) {
  gns::Params p;
  if (gns::plan(p, B, HW, C, G) != 0 || gns::smem_bytes(p) > gns::kSmemLimit) return PSG_ERR_UNSUPPORTED;
  if (workspace_floats < (long long)(gns::off_ready(p) + (size_t)B)) return PSG_ERR_UNSUPPORTED;
  p.dy = (const __nv_bfloat16*)dy; p.lddy = ld_dy;
  p.x = (const __nv_bfloat16*)x; p.ldx = ld_x;
  p.dx = (__nv_bfloat16*)dx; p.lddx = ld_dx;
  p.gamma = gamma; p.beta = beta; p.stats = stats;
  p.partial = workspace;
  p.part = workspace + gns::off_part(p);
  p.ready = reinterpret_cast<int*>(workspace + gns::off_ready(p));
  p.colsum_out = dx_colsum; p.ld_colsum = ld_colsum;
  if (act) return accumulate_dx ? gns::run<true, true>(p, stream) : gns::run<true, false>(p, stream);
  return accumulate_dx ? gns::run<false, true>(p, stream) : gns::run<false, false>(p, stream);
}

int gns_timeout_flag() {
  int v = 0, zero = 0;
  cudaMemcpyFromSymbol(&v, gns::g_timeout_flag, sizeof(int));
  cudaMemcpyToSymbol(gns::g_timeout_flag, &zero, sizeof(int));
  return v;
}
