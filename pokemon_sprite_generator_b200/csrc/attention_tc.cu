// Tensor-core attention core for the bf16 path: batched small GEMMs (mma.sync m16n8k16, bf16 in / fp32 accumulate)
// plus row-softmax forward / backward kernels.  SURVEY.md §2.1 K6/K7.
//
// The attention problems of this U-Net are tiny (Lq in {196,49,16}, Lk in {196,49,16,text<=256}, head_dim 80..320), so
// the [B,h,Lq,Lk] score matrix is cheap to materialise (<= 160 MB transient) and every product becomes a batched GEMM
// over (batch, head) with the four operand orientations below.  tcgen05 tiles (M=128) would be mostly padding at these
// sizes, so this op family uses the warp-level tensor-core path instead; the projections around it are tcgen05.
//
//   forward : S = scale * Q K^T            (A normal,  B normal)     -> softmax (+dropout) -> P, Pd
//             O = Pd V                     (A normal,  B transposed)
//   backward: dPd = dO V^T                 (A normal,  B normal)     -> dS = P o (mask*dPd/(1-p) - D)
//             dV  = Pd^T dO                (A transposed, B transposed)
//             dQ  = scale * dS K           (A normal,  B transposed)
//             dK  = scale * dS^T Q         (A transposed, B transposed)
//
// Reference semantics: nn.MultiheadAttention core, src/models/unet.py:160-173,217,235 (dropout on probabilities).
#include "psg_common.cuh"

namespace bmm {

constexpr int BM = 64, BN = 64, BK = 32;
constexpr int kThreads = 128;
constexpr int PAD = 8;

struct Mat {
  const __nv_bfloat16* ptr;
  long long sb, sh;   // element strides between batches / heads
  long long ld;       // row pitch of the stored matrix
  int trans;          // 0: stored [rows(M|N)][K]; 1: stored [K][rows]
};

struct Params {
  Mat a, b;
  void* c;
  long long c_sb, c_sh, ldc;
  int c_f32;
  int M, N, K, H;
  float alpha;
};

__device__ __forceinline__ void ldsm_x4(uint32_t (&r)[4], const void* p) {
  uint32_t a = (uint32_t)__cvta_generic_to_shared(p);
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.shared.b16 {%0,%1,%2,%3}, [%4];" : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]) : "r"(a));
}
__device__ __forceinline__ void ldsm_x4_t(uint32_t (&r)[4], const void* p) {
  uint32_t a = (uint32_t)__cvta_generic_to_shared(p);
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.trans.shared.b16 {%0,%1,%2,%3}, [%4];" : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]) : "r"(a));
}
__device__ __forceinline__ void mma16816(float (&d)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
  asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
               : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3])
               : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}

// Load a [R rows x Cc contiguous] bf16 tile (zero-filled outside [rmax, cmax]) into smem with row pitch Cc+PAD.
template <int R, int Cc>
__device__ __forceinline__ void load_tile(__nv_bfloat16* dst, const __nv_bfloat16* src, long long ld, int r0, int c0, int rmax, int cmax) {
  constexpr int VPR = Cc / 8;
  for (int idx = threadIdx.x; idx < R * VPR; idx += kThreads) {
    const int r = idx / VPR, v = idx - r * VPR;
    const int gr = r0 + r, gc = c0 + v * 8;
    uint4 val = make_uint4(0u, 0u, 0u, 0u);
    if (gr < rmax && gc < cmax) {
      const __nv_bfloat16* p = src + (long long)gr * ld + gc;
      if (gc + 8 <= cmax && ((reinterpret_cast<uintptr_t>(p) & 15) == 0)) {
        val = *reinterpret_cast<const uint4*>(p);
      } else {
        __nv_bfloat16 tmp[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) tmp[j] = (gc + j < cmax) ? p[j] : __float2bfloat16(0.f);
        val = *reinterpret_cast<uint4*>(tmp);
      }
    }
    *reinterpret_cast<uint4*>(dst + r * (Cc + PAD) + v * 8) = val;
  }
}

template <int TA, int TB>
__global__ void __launch_bounds__(kThreads) bmm_kernel(Params p) {
  // A tile: !TA -> [BM][BK] (k contiguous) ; TA -> [BK][BM] (m contiguous).  Same for B with BN.
  __shared__ __align__(16) __nv_bfloat16 sA[TA ? BK * (BM + PAD) : BM * (BK + PAD)];
  __shared__ __align__(16) __nv_bfloat16 sB[TB ? BK * (BN + PAD) : BN * (BK + PAD)];
  const int z = blockIdx.z, b = z / p.H, h = z - b * p.H;
  const __nv_bfloat16* A = p.a.ptr + b * p.a.sb + h * p.a.sh;
  const __nv_bfloat16* Bm = p.b.ptr + b * p.b.sb + h * p.b.sh;
  const int m0 = blockIdx.y * BM, n0 = blockIdx.x * BN;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int wm = (warp >> 1) * 32, wn = (warp & 1) * 32;
  float acc[2][4][4];
#pragma unroll
  for (int i = 0; i < 2; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j)
#pragma unroll
      for (int e = 0; e < 4; ++e) acc[i][j][e] = 0.f;

  for (int k0 = 0; k0 < p.K; k0 += BK) {
    if (TA) load_tile<BK, BM>(sA, A, p.a.ld, k0, m0, p.K, p.M);
    else    load_tile<BM, BK>(sA, A, p.a.ld, m0, k0, p.M, p.K);
    if (TB) load_tile<BK, BN>(sB, Bm, p.b.ld, k0, n0, p.K, p.N);
    else    load_tile<BN, BK>(sB, Bm, p.b.ld, n0, k0, p.N, p.K);
    __syncthreads();
#pragma unroll
    for (int kk = 0; kk < BK; kk += 16) {
      uint32_t af[2][4];
#pragma unroll
      for (int i = 0; i < 2; ++i) {
        const int mt = wm + i * 16;
        const int j = lane >> 3, r = lane & 7;
        if (TA) {
          // matrices: (m 0-7,k 0-7) (m 8-15,k 0-7) (m 0-7,k 8-15) (m 8-15,k 8-15); stored [k][m]
          ldsm_x4_t(af[i], &sA[(kk + r + 8 * (j >> 1)) * (BM + PAD) + mt + 8 * (j & 1)]);
        } else {
          ldsm_x4(af[i], &sA[(mt + r + 8 * (j & 1)) * (BK + PAD) + kk + 8 * (j >> 1)]);
        }
      }
#pragma unroll
      for (int jn = 0; jn < 4; jn += 2) {
        // two n8 tiles per ldmatrix.x4: regs {b0,b1} of tile jn, {b0,b1} of tile jn+1
        uint32_t bf[4];
        const int nt = wn + jn * 8;
        const int j = lane >> 3, r = lane & 7;
        if (TB) {
          // stored [k][n]; matrices: (k 0-7,n 0-7) (k 8-15,n 0-7) (k 0-7,n 8-15) (k 8-15,n 8-15)
          ldsm_x4_t(bf, &sB[(kk + r + 8 * (j & 1)) * (BN + PAD) + nt + 8 * (j >> 1)]);
        } else {
          // stored [n][k]; matrices: (n 0-7,k 0-7) (n 0-7,k 8-15) (n 8-15,k 0-7) (n 8-15,k 8-15)
          ldsm_x4(bf, &sB[(nt + r + 8 * (j >> 1)) * (BK + PAD) + kk + 8 * (j & 1)]);
        }
#pragma unroll
        for (int i = 0; i < 2; ++i) {
          mma16816(acc[i][jn], af[i], bf[0], bf[1]);
          mma16816(acc[i][jn + 1], af[i], bf[2], bf[3]);
        }
      }
    }
    __syncthreads();
  }
  // epilogue
  const long long cbase = b * p.c_sb + h * p.c_sh;
#pragma unroll
  for (int i = 0; i < 2; ++i)
#pragma unroll
    for (int jn = 0; jn < 4; ++jn)
#pragma unroll
      for (int half = 0; half < 2; ++half) {
        const int m = m0 + wm + i * 16 + (lane >> 2) + half * 8;
        const int n = n0 + wn + jn * 8 + (lane & 3) * 2;
        if (m >= p.M || n >= p.N) continue;
        const float v0 = acc[i][jn][half * 2] * p.alpha, v1 = acc[i][jn][half * 2 + 1] * p.alpha;
        const long long o = cbase + (long long)m * p.ldc + n;
        if (p.c_f32) {
          float* c = reinterpret_cast<float*>(p.c) + o;
          c[0] = v0;
          if (n + 1 < p.N) c[1] = v1;
        } else {
          __nv_bfloat16* c = reinterpret_cast<__nv_bfloat16*>(p.c) + o;
          if (n + 1 < p.N && ((o & 1) == 0)) *reinterpret_cast<__nv_bfloat162*>(c) = __floats2bfloat162_rn(v0, v1);
          else { c[0] = __float2bfloat16_rn(v0); if (n + 1 < p.N) c[1] = __float2bfloat16_rn(v1); }
        }
      }
}

// ---- row softmax over S [rows, ldp] fp32 -> P bf16 (and dropped copy Pd), one warp per row; Lk <= 1024 -------------
constexpr int kMaxPerLane = 32;

__global__ void softmax_fwd_kernel(const float* __restrict__ S, __nv_bfloat16* __restrict__ P, __nv_bfloat16* __restrict__ Pd,
                                   long long rows, int Lk, int ldp, unsigned long long seed, unsigned int thr, float keep_scale) {
  const long long row = (long long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (row >= rows) return;
  const int lane = threadIdx.x & 31;
  const float* s = S + row * ldp;
  float v[kMaxPerLane];
  float mx = -INFINITY;
#pragma unroll
  for (int i = 0; i < kMaxPerLane; ++i) {
    const int j = lane + 32 * i;
    v[i] = (j < Lk) ? s[j] : -INFINITY;
    mx = fmaxf(mx, v[i]);
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, o));
  float sum = 0.f;
#pragma unroll
  for (int i = 0; i < kMaxPerLane; ++i) {
    const int j = lane + 32 * i;
    v[i] = (j < Lk) ? __expf(v[i] - mx) : 0.f;
    sum += v[i];
  }
  sum = psg_warp_sum(sum);
  const float inv = 1.f / sum;
#pragma unroll
  for (int i = 0; i < kMaxPerLane; ++i) {
    const int j = lane + 32 * i;
    if (j >= ldp) break;
    const float pr = (j < Lk) ? v[i] * inv : 0.f;
    P[row * ldp + j] = __float2bfloat16_rn(pr);
    if (Pd != nullptr) {
      const bool kp = thr == 0 || psg_drop_keep(seed, (uint64_t)(row * Lk + j), thr);
      Pd[row * ldp + j] = __float2bfloat16_rn(kp ? pr * keep_scale : 0.f);
    }
  }
}

// dS = P o (m*dPd - D), D = sum_j P m dPd, m = mask/(1-p); also re-materialises Pd = P o m when dropout is on.
__global__ void softmax_bwd_kernel(const __nv_bfloat16* __restrict__ P, const float* __restrict__ dPd, __nv_bfloat16* __restrict__ dS,
                                   __nv_bfloat16* __restrict__ Pd, long long rows, int Lk, int ldp, unsigned long long seed,
                                   unsigned int thr, float keep_scale) {
  const long long row = (long long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (row >= rows) return;
  const int lane = threadIdx.x & 31;
  float pr[kMaxPerLane], g[kMaxPerLane];
  float d = 0.f;
#pragma unroll
  for (int i = 0; i < kMaxPerLane; ++i) {
    const int j = lane + 32 * i;
    pr[i] = 0.f; g[i] = 0.f;
    if (j < Lk) {
      pr[i] = __bfloat162float(P[row * ldp + j]);
      const bool kp = thr == 0 || psg_drop_keep(seed, (uint64_t)(row * Lk + j), thr);
      g[i] = kp ? dPd[row * ldp + j] * keep_scale : 0.f;
      if (Pd != nullptr) Pd[row * ldp + j] = __float2bfloat16_rn(kp ? pr[i] * keep_scale : 0.f);
      d = fmaf(pr[i], g[i], d);
    }
  }
  d = psg_warp_sum(d);
#pragma unroll
  for (int i = 0; i < kMaxPerLane; ++i) {
    const int j = lane + 32 * i;
    if (j >= ldp) break;
    dS[row * ldp + j] = __float2bfloat16_rn(j < Lk ? pr[i] * (g[i] - d) : 0.f);
    if (Pd != nullptr && j >= Lk) Pd[row * ldp + j] = __float2bfloat16(0.f);
  }
}

static int launch(const Params& p, int batches, cudaStream_t st) {
  dim3 grid((p.N + BN - 1) / BN, (p.M + BM - 1) / BM, batches);
  if (grid.y > 65535 || grid.z > 65535) { psg_set_error("psg_bmm_bf16: grid too large"); return PSG_ERR_INVALID; }
  const int ta = p.a.trans, tb = p.b.trans;
  if (!ta && !tb) bmm_kernel<0, 0><<<grid, kThreads, 0, st>>>(p);
  else if (!ta && tb) bmm_kernel<0, 1><<<grid, kThreads, 0, st>>>(p);
  else if (ta && !tb) bmm_kernel<1, 0><<<grid, kThreads, 0, st>>>(p);
  else bmm_kernel<1, 1><<<grid, kThreads, 0, st>>>(p);
  return PSG_OK;
}

}  // namespace bmm

extern "C" {

// C[b,h] (M x N) = alpha * A[b,h] (M x K) * B[b,h]^T (N x K); bf16 operands, fp32 accumulate, C bf16 or fp32.
// For X in {a, b}: base = ptr + b*sb + h*sh (elements); trans = 0 -> element (row, k) at base[row*ld + k];
// trans = 1 -> base[k*ld + row].  All pitches/offsets in elements.
int psg_bmm_bf16(const void* a, long long a_sb, long long a_sh, long long lda, int trans_a, const void* b, long long b_sb,
                 long long b_sh, long long ldb, int trans_b, void* c, long long c_sb, long long c_sh, long long ldc, int c_is_f32,
                 int batch, int heads, int M, int N, int K, float alpha, void* stream) {
  PSG_CHECK_ARG(a && b && c, "psg_bmm_bf16: null pointer");
  PSG_CHECK_ARG(batch > 0 && heads > 0 && M > 0 && N > 0 && K > 0, "psg_bmm_bf16: bad sizes");
  bmm::Params p;
  p.a = {(const __nv_bfloat16*)a, a_sb, a_sh, lda, trans_a};
  p.b = {(const __nv_bfloat16*)b, b_sb, b_sh, ldb, trans_b};
  p.c = c; p.c_sb = c_sb; p.c_sh = c_sh; p.ldc = ldc; p.c_f32 = c_is_f32;
  p.M = M; p.N = N; p.K = K; p.H = heads; p.alpha = alpha;
  int rc = bmm::launch(p, batch * heads, (cudaStream_t)stream);
  if (rc) return rc;
  PSG_CHECK_LAUNCH("psg_bmm_bf16");
  return PSG_OK;
}

// P = softmax(S) row-wise over Lk (S fp32 [rows, ldp], P bf16 [rows, ldp]); Pd = dropout(P) when Pd != null.
int psg_softmax_fwd(const float* S, void* P, void* Pd, long long rows, int Lk, int ldp, unsigned long long seed, float drop_p,
                    void* stream) {
  PSG_CHECK_ARG(S && P && rows > 0 && Lk > 0 && ldp >= Lk && ldp <= 32 * bmm::kMaxPerLane, "psg_softmax_fwd: bad args (Lk=%d ldp=%d)", Lk, ldp);
  const unsigned int thr = drop_p > 0.f ? (unsigned int)fmin((double)drop_p * 4294967296.0, 4294967295.0) : 0u;
  const float ks = drop_p > 0.f ? 1.f / (1.f - drop_p) : 1.f;
  const int wpb = 8;
  bmm::softmax_fwd_kernel<<<(unsigned)((rows + wpb - 1) / wpb), wpb * 32, 0, (cudaStream_t)stream>>>(
      S, (__nv_bfloat16*)P, (__nv_bfloat16*)Pd, rows, Lk, ldp, seed, thr, ks);
  PSG_CHECK_LAUNCH("psg_softmax_fwd");
  return PSG_OK;
}

int psg_softmax_bwd(const void* P, const float* dPd, void* dS, void* Pd, long long rows, int Lk, int ldp, unsigned long long seed,
                    float drop_p, void* stream) {
  PSG_CHECK_ARG(P && dPd && dS && rows > 0 && Lk > 0 && ldp >= Lk && ldp <= 32 * bmm::kMaxPerLane, "psg_softmax_bwd: bad args");
  const unsigned int thr = drop_p > 0.f ? (unsigned int)fmin((double)drop_p * 4294967296.0, 4294967295.0) : 0u;
  const float ks = drop_p > 0.f ? 1.f / (1.f - drop_p) : 1.f;
  const int wpb = 8;
  bmm::softmax_bwd_kernel<<<(unsigned)((rows + wpb - 1) / wpb), wpb * 32, 0, (cudaStream_t)stream>>>(
      (const __nv_bfloat16*)P, dPd, (__nv_bfloat16*)dS, (__nv_bfloat16*)Pd, rows, Lk, ldp, seed, thr, ks);
  PSG_CHECK_LAUNCH("psg_softmax_bwd");
  return PSG_OK;
}

}  // extern "C"
