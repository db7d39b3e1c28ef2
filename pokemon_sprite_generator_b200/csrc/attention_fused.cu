// Fused tensor-core attention (bf16 in, fp32 accumulate): softmax(scale * Q K^T) [dropout] V and its backward, with the
// score matrix kept on chip.  SURVEY.md §2.1 K6/K7; replaces the nn.MultiheadAttention core (src/models/unet.py:160-173,
// 217,235: scale 1/sqrt(head_dim), softmax over keys, dropout on the probabilities).
//
// The attention problems of this U-Net are small and numerous (B*heads up to 2048 problems of Lq in {196,49,16},
// Lk in {196,49,16, text tokens <= 256}, head_dim 80..320), so tcgen05 tiles (M = 128) would be mostly padding: each CTA
// takes one (batch, head, 64- or 32-row block), stages the whole K/V (or Q/dO) of that head in shared memory once and
// runs every product with warp-level mma.sync m16n8k16 out of shared memory.  Nothing of size Lq x Lk ever goes to
// global memory: the forward saves only the row log-sum-exp, the backward recomputes the probabilities.
//
//   forward  (q-block 64): S = scale Q K^T -> softmax -> LSE, Pd = drop(P) -> O = Pd V
//   backward (q-block 32): delta = rowsum(dO o O); S; dPd = dO V^T; dS = P o (drop(dPd) - delta); dQ = scale dS K
//   backward (k-block 32): S^T = scale K Q^T; dPd^T = V dO^T; Pd^T, dS^T; dV = Pd^T dO; dK = scale dS^T Q
//
// The dropout mask is the library-wide stateless rule psg_drop_keep(seed, ((b*H + h)*Lq + i)*Lk + j).
#include "psg_common.cuh"

namespace fattn {

constexpr int kThreads = 256;
constexpr int kWarps = kThreads / 32;
constexpr int kMaxLk = 256;          // softmax keeps a row in registers: Lk16 / 32 <= 8 values per lane
constexpr size_t kSmemLimit = 225 * 1024;

struct Params {
  const __nv_bfloat16 *q, *k, *v, *o, *dout;
  __nv_bfloat16 *out, *dq, *dk, *dv;
  float *lse, *delta;                 // [B, H, Lq]
  long long ldq, ldk, ldv, ldo, lddo, lddq, lddk, lddv;
  int B, H, Lq, Lk, hd;
  int hp;                             // smem row pitch (elements) of head_dim-wide tiles: hd + 8
  int Lk16, Lq16;                     // lengths rounded up to 16 (contraction padding)
  float scale;
  unsigned long long seed;
  unsigned int thr;
  float ks;                           // 1 / (1 - p)
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
__device__ __forceinline__ void cp_async16(void* smem_dst, const void* gmem_src, int src_bytes) {
  const uint32_t d = (uint32_t)__cvta_generic_to_shared(smem_dst);
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(d), "l"(gmem_src), "r"(src_bytes) : "memory");
}
__device__ __forceinline__ void cp_async_wait_all() { asm volatile("cp.async.commit_group;\n\tcp.async.wait_group 0;" ::: "memory"); }

// rows [r0, r0 + nrows) of a token-major [*, hd] matrix -> smem tile [rows_padded][hp]; rows >= rmax are zero-filled
__device__ __forceinline__ void load_rows(__nv_bfloat16* dst, int hp, const __nv_bfloat16* src, long long ld, int r0, int rmax,
                                          int rows_padded, int hd) {
  const int vpr = hd >> 3;
  for (int idx = threadIdx.x; idx < rows_padded * vpr; idx += kThreads) {
    const int r = idx / vpr, v = idx - r * vpr;
    const bool ok = r0 + r < rmax;
    cp_async16(dst + r * hp + v * 8, ok ? src + (long long)(r0 + r) * ld + v * 8 : src, ok ? 16 : 0);
  }
}

// One warp: acc[4][4] (16 x 32 outputs at rows m0.., columns n0..) = A[m][k] * B over k in [0, K), K % 16 == 0.
//   A stored [m][k] (pitch lda).  kBT == false: B stored [n][k] (pitch ldb); kBT == true: B stored [k][n].
template <bool kBT>
__device__ __forceinline__ void warp_mma_16x32(float (&acc)[4][4], const __nv_bfloat16* A, int lda, int m0, const __nv_bfloat16* Bm,
                                               int ldb, int n0, int K, int lane) {
#pragma unroll
  for (int j = 0; j < 4; ++j)
#pragma unroll
    for (int e = 0; e < 4; ++e) acc[j][e] = 0.f;
  const int j = lane >> 3, r = lane & 7;
  for (int kk = 0; kk < K; kk += 16) {
    uint32_t af[4];
    ldsm_x4(af, A + (m0 + r + 8 * (j & 1)) * lda + kk + 8 * (j >> 1));
#pragma unroll
    for (int jn = 0; jn < 4; jn += 2) {
      uint32_t bf[4];
      const int nt = n0 + jn * 8;
      if (kBT) ldsm_x4_t(bf, Bm + (kk + r + 8 * (j & 1)) * ldb + nt + 8 * (j >> 1));
      else     ldsm_x4(bf, Bm + (nt + r + 8 * (j >> 1)) * ldb + kk + 8 * (j & 1));
      mma16816(acc[jn], af, bf[0], bf[1]);
      mma16816(acc[jn + 1], af, bf[2], bf[3]);
    }
  }
}
// element (e) of acc[jn] sits at row m0 + (lane >> 2) + 8 * (e >> 1), column n0 + jn * 8 + (lane & 3) * 2 + (e & 1)

__device__ __forceinline__ bool keep_ij(const Params& p, long long row_global, int j) {
  return p.thr == 0 || psg_drop_keep(p.seed, (uint64_t)(row_global * p.Lk + j), p.thr);
}

// ---------------------------------------------------------------------------------------------------------------------
// forward: grid (ceil(Lq / 64), B * H)
//   smem: K [Lk16][hp] | V [Lk16][hp] | Q [64][hp] | S fp32 [64][sp]   (the bf16 Pd rows overwrite their own S rows)
// ---------------------------------------------------------------------------------------------------------------------
constexpr int kQB = 64;

__global__ void __launch_bounds__(kThreads) attn_fwd_kernel(const Params p) {
  extern __shared__ __align__(16) uint8_t smem[];
  const int bh = blockIdx.y, b = bh / p.H, h = bh - b * p.H;
  const int q0 = blockIdx.x * kQB;
  const int hp = p.hp, sp = p.Lk16 + 4;
  __nv_bfloat16* Ks = reinterpret_cast<__nv_bfloat16*>(smem);
  __nv_bfloat16* Vs = Ks + (size_t)p.Lk16 * hp;
  __nv_bfloat16* Qs = Vs + (size_t)p.Lk16 * hp;
  float* Sf = reinterpret_cast<float*>(Qs + (size_t)kQB * hp);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

  load_rows(Ks, hp, p.k + (long long)b * p.Lk * p.ldk + h * p.hd, p.ldk, 0, p.Lk, p.Lk16, p.hd);
  load_rows(Vs, hp, p.v + (long long)b * p.Lk * p.ldv + h * p.hd, p.ldv, 0, p.Lk, p.Lk16, p.hd);
  load_rows(Qs, hp, p.q + (long long)b * p.Lq * p.ldq + h * p.hd, p.ldq, q0, p.Lq, kQB, p.hd);
  cp_async_wait_all();
  __syncthreads();

  // S = scale * Q K^T
  const int nchunks = (p.Lk16 + 31) / 32;
  for (int item = warp; item < 4 * nchunks; item += kWarps) {
    const int mt = item & 3, nc = item >> 2;
    float acc[4][4];
    // the last chunk may run 16 columns past Lk16: those B rows belong to the V tile (finite data), results are dropped
    warp_mma_16x32<false>(acc, Qs, hp, mt * 16, Ks, hp, nc * 32, p.hd, lane);
#pragma unroll
    for (int jn = 0; jn < 4; ++jn)
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        const int m = mt * 16 + (lane >> 2) + 8 * (e >> 1), n = nc * 32 + jn * 8 + (lane & 3) * 2 + (e & 1);
        if (n < p.Lk16) Sf[m * sp + n] = acc[jn][e] * p.scale;
      }
  }
  __syncthreads();

  // row softmax (one warp per row), LSE out, dropout, bf16 Pd written over the row's own fp32 storage
  for (int r = warp; r < kQB; r += kWarps) {
    float v[kMaxLk / 32];
    float mx = -INFINITY;
    float* srow = Sf + r * sp;
#pragma unroll
    for (int i = 0; i < kMaxLk / 32; ++i) {
      const int jx = lane + 32 * i;
      v[i] = (jx < p.Lk) ? srow[jx] : -INFINITY;
      mx = fmaxf(mx, v[i]);
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, o));
    float sum = 0.f;
#pragma unroll
    for (int i = 0; i < kMaxLk / 32; ++i) {
      v[i] = (lane + 32 * i < p.Lk) ? __expf(v[i] - mx) : 0.f;
      sum += v[i];
    }
    sum = psg_warp_sum(sum);
    const float inv = 1.f / sum;
    const bool valid = q0 + r < p.Lq;
    const long long row_global = ((long long)bh) * p.Lq + q0 + r;
    if (valid && lane == 0 && p.lse) p.lse[row_global] = mx + __logf(sum);
    __syncwarp();
    __nv_bfloat16* prow = reinterpret_cast<__nv_bfloat16*>(srow);
#pragma unroll
    for (int i = 0; i < kMaxLk / 32; ++i) {
      const int jx = lane + 32 * i;
      if (jx < p.Lk16) {
        float pv = v[i] * inv;
        if (p.thr) pv = (valid && jx < p.Lk && keep_ij(p, row_global, jx)) ? pv * p.ks : 0.f;
        prow[jx] = __float2bfloat16_rn(valid ? pv : 0.f);
      }
    }
  }
  __syncthreads();

  // O = Pd V
  const __nv_bfloat16* Pd = reinterpret_cast<const __nv_bfloat16*>(Sf);
  const int ldp = 2 * sp;
  __nv_bfloat16* obase = p.out + (long long)b * p.Lq * p.ldo + h * p.hd;
  const int ochunks = (p.hd + 31) / 32;      // head_dim % 32 == 16: the last chunk's upper half is dropped
  for (int item = warp; item < 4 * ochunks; item += kWarps) {
    const int mt = item & 3, nc = item >> 2;
    float acc[4][4];
    warp_mma_16x32<true>(acc, Pd, ldp, mt * 16, Vs, hp, nc * 32, p.Lk16, lane);
#pragma unroll
    for (int jn = 0; jn < 4; ++jn)
#pragma unroll
      for (int half = 0; half < 2; ++half) {
        const int m = q0 + mt * 16 + (lane >> 2) + 8 * half, n = nc * 32 + jn * 8 + (lane & 3) * 2;
        if (m < p.Lq && n < p.hd) *reinterpret_cast<__nv_bfloat162*>(obase + (long long)m * p.ldo + n) = __floats2bfloat162_rn(acc[jn][2 * half], acc[jn][2 * half + 1]);
      }
  }
}

// ---------------------------------------------------------------------------------------------------------------------
// backward, dQ: grid (ceil(Lq / 32), B * H)
//   smem: K [Lk16][hp] | V [Lk16][hp] | Q [32][hp] | dO [32][hp] | S fp32 [32][sp] | dS bf16 [32][dp] | lse[32] | delta[32]
// ---------------------------------------------------------------------------------------------------------------------
constexpr int kBB = 32;

__global__ void __launch_bounds__(kThreads) attn_bwd_dq_kernel(const Params p) {
  extern __shared__ __align__(16) uint8_t smem[];
  const int bh = blockIdx.y, b = bh / p.H, h = bh - b * p.H;
  const int q0 = blockIdx.x * kBB;
  const int hp = p.hp, sp = p.Lk16 + 4, dp = p.Lk16 + 8;
  __nv_bfloat16* Ks = reinterpret_cast<__nv_bfloat16*>(smem);
  __nv_bfloat16* Vs = Ks + (size_t)p.Lk16 * hp;
  __nv_bfloat16* Qs = Vs + (size_t)p.Lk16 * hp;
  __nv_bfloat16* dOs = Qs + (size_t)kBB * hp;
  float* Sf = reinterpret_cast<float*>(dOs + (size_t)kBB * hp);
  __nv_bfloat16* dSs = reinterpret_cast<__nv_bfloat16*>(Sf + (size_t)kBB * sp);
  float* lse_s = reinterpret_cast<float*>(dSs + (size_t)kBB * dp);
  float* del_s = lse_s + kBB;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

  load_rows(Ks, hp, p.k + (long long)b * p.Lk * p.ldk + h * p.hd, p.ldk, 0, p.Lk, p.Lk16, p.hd);
  load_rows(Vs, hp, p.v + (long long)b * p.Lk * p.ldv + h * p.hd, p.ldv, 0, p.Lk, p.Lk16, p.hd);
  load_rows(Qs, hp, p.q + (long long)b * p.Lq * p.ldq + h * p.hd, p.ldq, q0, p.Lq, kBB, p.hd);
  load_rows(dOs, hp, p.dout + (long long)b * p.Lq * p.lddo + h * p.hd, p.lddo, q0, p.Lq, kBB, p.hd);
  // delta[i] = dO_i . O_i (O straight from global), one warp per row; also stage lse
  for (int r = warp; r < kBB; r += kWarps) {
    const int qi = q0 + r;
    float s = 0.f;
    if (qi < p.Lq) {
      const __nv_bfloat16* orow = p.o + ((long long)b * p.Lq + qi) * p.ldo + h * p.hd;
      const __nv_bfloat16* drow = p.dout + ((long long)b * p.Lq + qi) * p.lddo + h * p.hd;
      for (int c = lane * 8; c < p.hd; c += 256) {
        Vec8<__nv_bfloat16> a, d;
        a.load(orow + c);
        d.load(drow + c);
#pragma unroll
        for (int x = 0; x < 8; ++x) s = fmaf(a.v[x], d.v[x], s);
      }
    }
    s = psg_warp_sum(s);
    if (lane == 0) {
      const long long rg = (long long)bh * p.Lq + qi;
      del_s[r] = s;
      lse_s[r] = (qi < p.Lq) ? p.lse[rg] : 0.f;
      if (qi < p.Lq) p.delta[rg] = s;
    }
  }
  cp_async_wait_all();
  __syncthreads();

  const int nchunks = (p.Lk16 + 31) / 32;
  // S = scale * Q K^T -> smem
  for (int item = warp; item < 2 * nchunks; item += kWarps) {
    const int mt = item & 1, nc = item >> 1;
    float acc[4][4];
    warp_mma_16x32<false>(acc, Qs, hp, mt * 16, Ks, hp, nc * 32, p.hd, lane);
#pragma unroll
    for (int jn = 0; jn < 4; ++jn)
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        const int m = mt * 16 + (lane >> 2) + 8 * (e >> 1), n = nc * 32 + jn * 8 + (lane & 3) * 2 + (e & 1);
        if (n < p.Lk16) Sf[m * sp + n] = acc[jn][e] * p.scale;
      }
  }
  __syncthreads();
  // dPd = dO V^T ; dS = P o (drop(dPd) - delta)
  for (int item = warp; item < 2 * nchunks; item += kWarps) {
    const int mt = item & 1, nc = item >> 1;
    float acc[4][4];
    warp_mma_16x32<false>(acc, dOs, hp, mt * 16, Vs, hp, nc * 32, p.hd, lane);
#pragma unroll
    for (int jn = 0; jn < 4; ++jn)
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        const int m = mt * 16 + (lane >> 2) + 8 * (e >> 1), n = nc * 32 + jn * 8 + (lane & 3) * 2 + (e & 1);
        if (n >= p.Lk16) continue;
        float ds = 0.f;
        if (n < p.Lk && q0 + m < p.Lq) {
          const float pr = __expf(Sf[m * sp + n] - lse_s[m]);
          float dpv = acc[jn][e];
          if (p.thr) dpv = keep_ij(p, (long long)bh * p.Lq + q0 + m, n) ? dpv * p.ks : 0.f;
          ds = pr * (dpv - del_s[m]);
        }
        dSs[m * dp + n] = __float2bfloat16_rn(ds);
      }
  }
  __syncthreads();
  // dQ = scale * dS K
  __nv_bfloat16* qbase = p.dq + (long long)b * p.Lq * p.lddq + h * p.hd;
  const int ochunks = (p.hd + 31) / 32;      // head_dim % 32 == 16: the last chunk's upper half is dropped
  for (int item = warp; item < 2 * ochunks; item += kWarps) {
    const int mt = item & 1, nc = item >> 1;
    float acc[4][4];
    warp_mma_16x32<true>(acc, dSs, dp, mt * 16, Ks, hp, nc * 32, p.Lk16, lane);
#pragma unroll
    for (int jn = 0; jn < 4; ++jn)
#pragma unroll
      for (int half = 0; half < 2; ++half) {
        const int m = q0 + mt * 16 + (lane >> 2) + 8 * half, n = nc * 32 + jn * 8 + (lane & 3) * 2;
        if (m < p.Lq && n < p.hd)
          *reinterpret_cast<__nv_bfloat162*>(qbase + (long long)m * p.lddq + n) =
              __floats2bfloat162_rn(acc[jn][2 * half] * p.scale, acc[jn][2 * half + 1] * p.scale);
      }
  }
}

// ---------------------------------------------------------------------------------------------------------------------
// backward, dK / dV: grid (ceil(Lk / 32), B * H)
//   smem: Q [Lq16][hp] | dO [Lq16][hp] | K [32][hp] | V [32][hp] | S^T fp32 [32][sp] | Pd^T bf16 [32][dp] | dS^T bf16 [32][dp]
//         | lse[Lq16] | delta[Lq16]
// ---------------------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(kThreads) attn_bwd_dkv_kernel(const Params p) {
  extern __shared__ __align__(16) uint8_t smem[];
  const int bh = blockIdx.y, b = bh / p.H, h = bh - b * p.H;
  const int k0 = blockIdx.x * kBB;
  const int hp = p.hp, sp = p.Lq16 + 4, dp = p.Lq16 + 8;
  __nv_bfloat16* Qs = reinterpret_cast<__nv_bfloat16*>(smem);
  __nv_bfloat16* dOs = Qs + (size_t)p.Lq16 * hp;
  __nv_bfloat16* Ks = dOs + (size_t)p.Lq16 * hp;
  __nv_bfloat16* Vs = Ks + (size_t)kBB * hp;
  float* Sf = reinterpret_cast<float*>(Vs + (size_t)kBB * hp);
  __nv_bfloat16* Pt = reinterpret_cast<__nv_bfloat16*>(Sf + (size_t)kBB * sp);
  __nv_bfloat16* dSt = Pt + (size_t)kBB * dp;
  float* lse_s = reinterpret_cast<float*>(dSt + (size_t)kBB * dp);
  float* del_s = lse_s + p.Lq16;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

  load_rows(Qs, hp, p.q + (long long)b * p.Lq * p.ldq + h * p.hd, p.ldq, 0, p.Lq, p.Lq16, p.hd);
  load_rows(dOs, hp, p.dout + (long long)b * p.Lq * p.lddo + h * p.hd, p.lddo, 0, p.Lq, p.Lq16, p.hd);
  load_rows(Ks, hp, p.k + (long long)b * p.Lk * p.ldk + h * p.hd, p.ldk, k0, p.Lk, kBB, p.hd);
  load_rows(Vs, hp, p.v + (long long)b * p.Lk * p.ldv + h * p.hd, p.ldv, k0, p.Lk, kBB, p.hd);
  for (int i = threadIdx.x; i < p.Lq16; i += kThreads) {
    const bool ok = i < p.Lq;
    lse_s[i] = ok ? p.lse[(long long)bh * p.Lq + i] : 0.f;
    del_s[i] = ok ? p.delta[(long long)bh * p.Lq + i] : 0.f;
  }
  cp_async_wait_all();
  __syncthreads();

  const int nchunks = (p.Lq16 + 31) / 32;
  // S^T = scale * K Q^T (rows = keys of this block, columns = queries)
  for (int item = warp; item < 2 * nchunks; item += kWarps) {
    const int mt = item & 1, nc = item >> 1;
    float acc[4][4];
    warp_mma_16x32<false>(acc, Ks, hp, mt * 16, Qs, hp, nc * 32, p.hd, lane);
#pragma unroll
    for (int jn = 0; jn < 4; ++jn)
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        const int m = mt * 16 + (lane >> 2) + 8 * (e >> 1), n = nc * 32 + jn * 8 + (lane & 3) * 2 + (e & 1);
        if (n < p.Lq16) Sf[m * sp + n] = acc[jn][e] * p.scale;
      }
  }
  __syncthreads();
  // dPd^T = V dO^T ; Pd^T and dS^T
  for (int item = warp; item < 2 * nchunks; item += kWarps) {
    const int mt = item & 1, nc = item >> 1;
    float acc[4][4];
    warp_mma_16x32<false>(acc, Vs, hp, mt * 16, dOs, hp, nc * 32, p.hd, lane);
#pragma unroll
    for (int jn = 0; jn < 4; ++jn)
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        const int m = mt * 16 + (lane >> 2) + 8 * (e >> 1), n = nc * 32 + jn * 8 + (lane & 3) * 2 + (e & 1);
        if (n >= p.Lq16) continue;
        float pd = 0.f, ds = 0.f;
        if (n < p.Lq && k0 + m < p.Lk) {
          const float pr = __expf(Sf[m * sp + n] - lse_s[n]);
          float dpv = acc[jn][e];
          pd = pr;
          if (p.thr) {
            const bool kp = keep_ij(p, (long long)bh * p.Lq + n, k0 + m);
            pd = kp ? pr * p.ks : 0.f;
            dpv = kp ? dpv * p.ks : 0.f;
          }
          ds = pr * (dpv - del_s[n]);
        }
        Pt[m * dp + n] = __float2bfloat16_rn(pd);
        dSt[m * dp + n] = __float2bfloat16_rn(ds);
      }
  }
  __syncthreads();
  // dV = Pd^T dO ; dK = scale * dS^T Q
  __nv_bfloat16* vbase = p.dv + (long long)b * p.Lk * p.lddv + h * p.hd;
  __nv_bfloat16* kbase = p.dk + (long long)b * p.Lk * p.lddk + h * p.hd;
  const int ochunks = (p.hd + 31) / 32;      // head_dim % 32 == 16: the last chunk's upper half is dropped
  for (int item = warp; item < 4 * ochunks; item += kWarps) {
    const int which = item & 1, mt = (item >> 1) & 1, nc = item >> 2;
    float acc[4][4];
    if (which == 0) warp_mma_16x32<true>(acc, Pt, dp, mt * 16, dOs, hp, nc * 32, p.Lq16, lane);
    else            warp_mma_16x32<true>(acc, dSt, dp, mt * 16, Qs, hp, nc * 32, p.Lq16, lane);
    const float sc = which == 0 ? 1.f : p.scale;
    __nv_bfloat16* base = which == 0 ? vbase : kbase;
    const long long ld = which == 0 ? p.lddv : p.lddk;
#pragma unroll
    for (int jn = 0; jn < 4; ++jn)
#pragma unroll
      for (int half = 0; half < 2; ++half) {
        const int m = k0 + mt * 16 + (lane >> 2) + 8 * half, n = nc * 32 + jn * 8 + (lane & 3) * 2;
        if (m < p.Lk && n < p.hd)
          *reinterpret_cast<__nv_bfloat162*>(base + (long long)m * ld + n) = __floats2bfloat162_rn(acc[jn][2 * half] * sc, acc[jn][2 * half + 1] * sc);
      }
  }
}

static size_t fwd_smem(const Params& p) {
  return ((size_t)2 * p.Lk16 * p.hp + (size_t)kQB * p.hp) * 2 + (size_t)kQB * (p.Lk16 + 4) * 4 + 64;
}
static size_t dq_smem(const Params& p) {
  return ((size_t)2 * p.Lk16 * p.hp + (size_t)2 * kBB * p.hp) * 2 + (size_t)kBB * (p.Lk16 + 4) * 4 + (size_t)kBB * (p.Lk16 + 8) * 2 +
         2 * kBB * 4 + 64;
}
static size_t dkv_smem(const Params& p) {
  return ((size_t)2 * p.Lq16 * p.hp + (size_t)2 * kBB * p.hp) * 2 + (size_t)kBB * (p.Lq16 + 4) * 4 + (size_t)2 * kBB * (p.Lq16 + 8) * 2 +
         (size_t)2 * p.Lq16 * 4 + 64;
}

static int fill(Params& p, int B, int H, int Lq, int Lk, int hd, float scale, unsigned long long seed, float drop_p) {
  if (B <= 0 || H <= 0 || Lq <= 0 || Lk <= 0 || hd <= 0 || hd % 16 != 0 || Lk > kMaxLk || Lq > 1024) return -1;
  if ((long long)B * H > 65535) return -1;
  p.B = B; p.H = H; p.Lq = Lq; p.Lk = Lk; p.hd = hd;
  p.hp = hd + 8;
  p.Lk16 = (Lk + 15) / 16 * 16;
  p.Lq16 = (Lq + 15) / 16 * 16;
  p.scale = scale;
  p.seed = seed;
  p.thr = drop_p > 0.f ? (unsigned int)fmin((double)drop_p * 4294967296.0, 4294967295.0) : 0u;
  p.ks = drop_p > 0.f ? 1.f / (1.f - drop_p) : 1.f;
  return 0;
}

template <typename Kern>
static int configure(Kern kern, bool& done, const char* name) {
  if (done) return PSG_OK;
  cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kSmemLimit);
  if (e != cudaSuccess) { psg_set_error("%s: cudaFuncSetAttribute: %s", name, cudaGetErrorString(e)); return PSG_ERR_CUDA; }
  done = true;
  return PSG_OK;
}

}  // namespace fattn

extern "C" {

// 1 if the fused kernels take this problem (bf16; head_dim % 16 == 0; Lk <= 256; everything fits in shared memory).
int psg_attn_fused_ok(int B, int H, int Lq, int Lk, int hd) {
  fattn::Params p;
  if (fattn::fill(p, B, H, Lq, Lk, hd, 1.f, 0, 0.f) != 0) return 0;
  // the S = Q K^T / dPd = dO V^T chunk loops may read 16 rows past the K (V, Q, dO) tile: the next tile must exist (it does)
  return fattn::fwd_smem(p) <= fattn::kSmemLimit && fattn::dq_smem(p) <= fattn::kSmemLimit && fattn::dkv_smem(p) <= fattn::kSmemLimit;
}

// o = softmax(scale q k^T) [dropout] v per (batch, head); q/o: [B*Lq, ld] (+ head*hd columns), k/v: [B*Lk, ld]; lse [B,H,Lq].
int psg_attn_fused_fwd(const void* q, long long ldq, const void* k, long long ldk, const void* v, long long ldv, void* o, long long ldo,
                       float* lse, int B, int H, int Lq, int Lk, int hd, float scale, unsigned long long drop_seed, float drop_p,
                       void* stream) {
  using namespace fattn;
  PSG_CHECK_ARG(q && k && v && o, "psg_attn_fused_fwd: null pointer");
  Params p;
  memset(&p, 0, sizeof(p));
  PSG_CHECK_ARG(fill(p, B, H, Lq, Lk, hd, scale, drop_seed, drop_p) == 0 && fwd_smem(p) <= kSmemLimit,
                "psg_attn_fused_fwd: unsupported problem B=%d H=%d Lq=%d Lk=%d hd=%d", B, H, Lq, Lk, hd);
  PSG_CHECK_ARG(ldq % 8 == 0 && ldk % 8 == 0 && ldv % 8 == 0 && ldo % 2 == 0 && ((uintptr_t)q % 16 == 0) && ((uintptr_t)k % 16 == 0) &&
                    ((uintptr_t)v % 16 == 0) && ((uintptr_t)o % 4 == 0),
                "psg_attn_fused_fwd: pitches/pointers must be 16B aligned");
  p.q = (const __nv_bfloat16*)q; p.k = (const __nv_bfloat16*)k; p.v = (const __nv_bfloat16*)v; p.out = (__nv_bfloat16*)o;
  p.ldq = ldq; p.ldk = ldk; p.ldv = ldv; p.ldo = ldo; p.lse = lse;
  static bool done = false;
  int rc = configure(attn_fwd_kernel, done, "psg_attn_fused_fwd");
  if (rc) return rc;
  dim3 grid((Lq + kQB - 1) / kQB, B * H);
  attn_fwd_kernel<<<grid, kThreads, fwd_smem(p), (cudaStream_t)stream>>>(p);
  PSG_CHECK_LAUNCH("psg_attn_fused_fwd");
  return PSG_OK;
}

// dq/dk/dv of the same problem; delta [B,H,Lq] is scratch (written by the dQ pass, read by the dK/dV pass).
int psg_attn_fused_bwd(const void* q, long long ldq, const void* k, long long ldk, const void* v, long long ldv, const void* o,
                       long long ldo, const void* dout, long long lddo, const float* lse, float* delta, void* dq, long long lddq,
                       void* dk, long long lddk, void* dv, long long lddv, int B, int H, int Lq, int Lk, int hd, float scale,
                       unsigned long long drop_seed, float drop_p, void* stream) {
  using namespace fattn;
  PSG_CHECK_ARG(q && k && v && o && dout && lse && delta && dq && dk && dv, "psg_attn_fused_bwd: null pointer");
  Params p;
  memset(&p, 0, sizeof(p));
  PSG_CHECK_ARG(fill(p, B, H, Lq, Lk, hd, scale, drop_seed, drop_p) == 0 && dq_smem(p) <= kSmemLimit && dkv_smem(p) <= kSmemLimit,
                "psg_attn_fused_bwd: unsupported problem B=%d H=%d Lq=%d Lk=%d hd=%d", B, H, Lq, Lk, hd);
  PSG_CHECK_ARG(ldq % 8 == 0 && ldk % 8 == 0 && ldv % 8 == 0 && ldo % 8 == 0 && lddo % 8 == 0 && lddq % 2 == 0 && lddk % 2 == 0 &&
                    lddv % 2 == 0 && ((uintptr_t)q % 16 == 0) && ((uintptr_t)k % 16 == 0) && ((uintptr_t)v % 16 == 0) &&
                    ((uintptr_t)o % 16 == 0) && ((uintptr_t)dout % 16 == 0),
                "psg_attn_fused_bwd: pitches/pointers must be 16B aligned");
  p.q = (const __nv_bfloat16*)q; p.k = (const __nv_bfloat16*)k; p.v = (const __nv_bfloat16*)v; p.o = (const __nv_bfloat16*)o;
  p.dout = (const __nv_bfloat16*)dout; p.dq = (__nv_bfloat16*)dq; p.dk = (__nv_bfloat16*)dk; p.dv = (__nv_bfloat16*)dv;
  p.ldq = ldq; p.ldk = ldk; p.ldv = ldv; p.ldo = ldo; p.lddo = lddo; p.lddq = lddq; p.lddk = lddk; p.lddv = lddv;
  p.lse = const_cast<float*>(lse); p.delta = delta;
  static bool done1 = false, done2 = false;
  int rc = configure(attn_bwd_dq_kernel, done1, "psg_attn_fused_bwd");
  if (rc) return rc;
  rc = configure(attn_bwd_dkv_kernel, done2, "psg_attn_fused_bwd");
  if (rc) return rc;
  cudaStream_t st = (cudaStream_t)stream;
  attn_bwd_dq_kernel<<<dim3((Lq + kBB - 1) / kBB, B * H), kThreads, dq_smem(p), st>>>(p);
  attn_bwd_dkv_kernel<<<dim3((Lk + kBB - 1) / kBB, B * H), kThreads, dkv_smem(p), st>>>(p);
  PSG_CHECK_LAUNCH("psg_attn_fused_bwd");
  g_psg_launch_count += 1;  // two kernels
  return PSG_OK;
}

}  // extern "C"
