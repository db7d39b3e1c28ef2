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
//   backward (q-block 64): delta = rowsum(dO o O); {S, dPd = dO V^T} per tile in registers -> dS = P o (drop(dPd) - delta)
//                          -> shared memory (bf16) -> dQ = scale dS K
//   backward (k-block 64): S^T -> Pd^T (bf16) -> dV = Pd^T dO; then {S^T, dPd^T = V dO^T} -> dS^T (same buffer) -> dK = scale dS^T Q
//                          (S^T is computed twice: FLOPs are free here, shared memory is not)
//
// The dropout mask is the library-wide stateless rule psg_drop_keep(seed, ((b*H + h)*Lq + i)*Lk + j).
#include "psg_common.cuh"

namespace fattn {

constexpr int kThreads = 512;      // big problems (one CTA per SM by shared memory): 16 warps; small ones launch 256 threads
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
  int bar_off;                        // byte offset of the two mbarriers in dynamic shared memory (end of the kernel's layout)
};

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void ldsm_x4(uint32_t (&r)[4], uint32_t a) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.shared.b16 {%0,%1,%2,%3}, [%4];" : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]) : "r"(a));
}
__device__ __forceinline__ void ldsm_x4_t(uint32_t (&r)[4], uint32_t a) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.trans.shared.b16 {%0,%1,%2,%3}, [%4];" : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]) : "r"(a));
}
__device__ __forceinline__ void mma16816(float (&d)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
  asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
               : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3])
               : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}

// ---- mbarrier + bulk-copy helpers: one cp.async.bulk per row (the rows of a head are hd*2 contiguous bytes in global
// memory and land at the padded pitch hp in shared memory).  The per-16-byte cp.async loops this replaces were 27-37 % of
// all executed instructions of these kernels (profiles/r01_ncu_attention_v2.md).
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint32_t bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t}"
      : "=r"(ok)
      : "r"(bar), "r"(parity)
      : "memory");
  return ok != 0;
}
// Bounded wait: a protocol bug must never hang the GPU; on timeout fall through (the parity tests then fail loudly).
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
  for (uint32_t i = 0; i < (1u << 24); ++i)
    if (mbar_try_wait(bar, parity)) return;
}
__device__ __forceinline__ void bulk_g2s(uint32_t dst, const void* src, uint32_t bytes, uint32_t bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst), "l"(src), "r"(bytes),
               "r"(bar)
               : "memory");
}
// bytes the live rows of a tile will deliver
__device__ __forceinline__ uint32_t tile_bytes(int r0, int rmax, int rows_padded, int hd) {
  const int live = max(0, min(rows_padded, rmax - r0));
  return (uint32_t)live * (uint32_t)hd * 2u;
}
// rows [r0, r0 + rows_padded) of a token-major [*, hd] matrix -> smem tile [rows_padded][hp]: one bulk copy per live row
// (completing on `bar`, whose expect_tx the caller posts with tile_bytes), plain zero stores for rows >= rmax.
__device__ __forceinline__ void load_rows(__nv_bfloat16* dst, int hp, const __nv_bfloat16* src, long long ld, int r0, int rmax,
                                          int rows_padded, int hd, uint32_t bar) {
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");      // earlier generic-proxy accesses of the tile vs the async writes
  for (int r = threadIdx.x; r < rows_padded; r += blockDim.x) {
    __nv_bfloat16* drow = dst + r * hp;
    if (r0 + r < rmax) {
      bulk_g2s(smem_u32(drow), src + (long long)(r0 + r) * ld, (uint32_t)hd * 2u, bar);
    } else {
      for (int v = 0; v < hd; v += 8) *reinterpret_cast<uint4*>(drow + v) = make_uint4(0u, 0u, 0u, 0u);
    }
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
  // shared-space byte addresses of this lane's ldmatrix rows, advanced by a constant per 16-deep k step: the loop body
  // is 3 LDSM + 4 HMMA + 3 adds (recomputing generic addresses per step made the integer pipe, not the tensor pipe, the limit)
  uint32_t a_addr = smem_u32(A + (m0 + r + 8 * (j & 1)) * lda + 8 * (j >> 1));
  uint32_t b_addr0, b_addr1, b_step;
  if (kBT) {
    b_addr0 = smem_u32(Bm + (r + 8 * (j & 1)) * ldb + n0 + 8 * (j >> 1));
    b_addr1 = b_addr0 + 32;                 // n0 + 16
    b_step = (uint32_t)ldb * 32u;           // 16 k rows
  } else {
    b_addr0 = smem_u32(Bm + (n0 + r + 8 * (j >> 1)) * ldb + 8 * (j & 1));
    b_addr1 = b_addr0 + (uint32_t)ldb * 32u;  // n0 + 16
    b_step = 32u;                           // 16 k elements
  }
#pragma unroll 2
  for (int kk = 0; kk < K; kk += 16) {
    uint32_t af[4], b0[4], b1[4];
    ldsm_x4(af, a_addr);
    if (kBT) { ldsm_x4_t(b0, b_addr0); ldsm_x4_t(b1, b_addr1); }
    else     { ldsm_x4(b0, b_addr0);   ldsm_x4(b1, b_addr1); }
    a_addr += 32u;
    b_addr0 += b_step;
    b_addr1 += b_step;
    mma16816(acc[0], af, b0[0], b0[1]);
    mma16816(acc[1], af, b0[2], b0[3]);
    mma16816(acc[2], af, b1[0], b1[1]);
    mma16816(acc[3], af, b1[2], b1[3]);
  }
}
// element (e) of acc[jn] sits at row m0 + (lane >> 2) + 8 * (e >> 1), column n0 + jn * 8 + (lane & 3) * 2 + (e & 1)

// Same product with BOTH operands stored reduction-major: A stored [k][m] (pitch lda: the m16k16 fragment comes out of
// ldmatrix.trans), B stored [k][n] (pitch ldb).  acc = A^T-stored[m0.., :] * B[:, n0..] over k in [0, K), K % 16 == 0.
__device__ __forceinline__ void warp_mma_16x32_at(float (&acc)[4][4], const __nv_bfloat16* At, int lda, int m0, const __nv_bfloat16* Bm,
                                                  int ldb, int n0, int K, int lane) {
#pragma unroll
  for (int j = 0; j < 4; ++j)
#pragma unroll
    for (int e = 0; e < 4; ++e) acc[j][e] = 0.f;
  const int j = lane >> 3, r = lane & 7;
  uint32_t a_addr = smem_u32(At + (r + 8 * (j >> 1)) * lda + m0 + 8 * (j & 1));
  uint32_t b_addr0 = smem_u32(Bm + (r + 8 * (j & 1)) * ldb + n0 + 8 * (j >> 1));
  uint32_t b_addr1 = b_addr0 + 32;
  const uint32_t a_step = (uint32_t)lda * 32u, b_step = (uint32_t)ldb * 32u;
#pragma unroll 2
  for (int kk = 0; kk < K; kk += 16) {
    uint32_t af[4], b0[4], b1[4];
    ldsm_x4_t(af, a_addr);
    ldsm_x4_t(b0, b_addr0);
    ldsm_x4_t(b1, b_addr1);
    a_addr += a_step;
    b_addr0 += b_step;
    b_addr1 += b_step;
    mma16816(acc[0], af, b0[0], b0[1]);
    mma16816(acc[1], af, b0[2], b0[3]);
    mma16816(acc[2], af, b1[0], b1[1]);
    mma16816(acc[3], af, b1[2], b1[3]);
  }
}

__device__ __forceinline__ bool keep_ij(const Params& p, long long row_global, int j) {
  return p.thr == 0 || psg_drop_keep(p.seed, (uint64_t)(row_global * p.Lk + j), p.thr);
}
// keep flags of the two neighbouring scores (row, j) and (row, j + 1) from at most two hashes (one when row*Lk + j is even)
__device__ __forceinline__ void keep_pair(const Params& p, uint64_t row_base, int j, bool& k0, bool& k1) {
  const uint64_t idx = row_base + (uint64_t)j;
  const uint32_t h0 = psg_hash32(p.seed, idx >> 1);
  k0 = psg_drop_keep2(h0, (int)(idx & 1), p.thr);
  const uint32_t h1 = (idx & 1) ? psg_hash32(p.seed, (idx + 1) >> 1) : h0;
  k1 = psg_drop_keep2(h1, (int)((idx + 1) & 1), p.thr);
}

// ---------------------------------------------------------------------------------------------------------------------
// All three kernels: grid (nsplit, B * H); a CTA keeps the head's resident tiles (K, V or Q, dO) in shared memory and walks
// the 64-row blocks blockIdx.x, blockIdx.x + nsplit, ... of that head, so the resident tiles are fetched once per head
// (nsplit == 1 whenever B * H alone fills the GPU) and the next block's tiles are fetched while the current block's last
// phase still runs.  bars[0]: resident tiles, bars[1]: per-block tiles (parity flips once per block).
// ---------------------------------------------------------------------------------------------------------------------
__device__ __forceinline__ void split_item(int item, int mtl, int& nc, int& mt) {     // item = nc * mtl + mt, mtl in 1..4
  nc = mtl == 4 ? item >> 2 : mtl == 1 ? item : mtl == 2 ? item >> 1 : item / 3;
  mt = item - nc * mtl;
}

// ---------------------------------------------------------------------------------------------------------------------
// forward
//   smem: K [Lk16][hp] | V [Lk16][hp] | Q [64][hp] | S fp32 [64][sp]   (the bf16 Pd rows overwrite their own S rows)
// ---------------------------------------------------------------------------------------------------------------------
constexpr int kQB = 64;

__global__ void __launch_bounds__(kThreads) attn_fwd_kernel(const Params p) {
  extern __shared__ __align__(16) uint8_t smem[];
  const int bh = blockIdx.y, b = bh / p.H, h = bh - b * p.H;
  const int hp = p.hp, sp = p.Lk16 + 4;
  __nv_bfloat16* Ks = reinterpret_cast<__nv_bfloat16*>(smem);
  __nv_bfloat16* Vs = Ks + (size_t)p.Lk16 * hp;
  __nv_bfloat16* Qs = Vs + (size_t)p.Lk16 * hp;
  float* Sf = reinterpret_cast<float*>(Qs + (size_t)kQB * hp);
  const uint32_t bar0 = smem_u32(smem + p.bar_off), bar1 = bar0 + 8;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, kWarps = blockDim.x >> 5;
  const int nqb = (p.Lq + kQB - 1) / kQB;
  const __nv_bfloat16* qsrc = p.q + (long long)b * p.Lq * p.ldq + h * p.hd;

  if (threadIdx.x == 0) {
    mbar_init(bar0, 1);
    mbar_init(bar1, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    mbar_expect_tx(bar0, 2u * tile_bytes(0, p.Lk, p.Lk16, p.hd));
    mbar_expect_tx(bar1, tile_bytes(blockIdx.x * kQB, p.Lq, kQB, p.hd));
  }
  __syncthreads();           // barriers initialised and their byte counts posted before any copy is issued
  load_rows(Ks, hp, p.k + (long long)b * p.Lk * p.ldk + h * p.hd, p.ldk, 0, p.Lk, p.Lk16, p.hd, bar0);
  load_rows(Qs, hp, qsrc, p.ldq, blockIdx.x * kQB, p.Lq, kQB, p.hd, bar1);
  load_rows(Vs, hp, p.v + (long long)b * p.Lk * p.ldv + h * p.hd, p.ldv, 0, p.Lk, p.Lk16, p.hd, bar0);
  mbar_wait(bar0, 0);

  uint32_t phase = 0;
  for (int qb = blockIdx.x; qb < nqb; qb += gridDim.x) {
    const int q0 = qb * kQB;
    const int mtl = min(kQB / 16, (p.Lq - q0 + 15) >> 4);      // live 16-row tiles of this block
    mbar_wait(bar1, phase);
    phase ^= 1;
    // the next block's byte count is posted a barrier ahead of its copies (issued after the S phase)
    if (threadIdx.x == 0 && qb + (int)gridDim.x < nqb) mbar_expect_tx(bar1, tile_bytes((qb + gridDim.x) * kQB, p.Lq, kQB, p.hd));
    __syncthreads();           // zero-filled rows are visible; the previous block's Pd reads are over

    // S = scale * Q K^T
    const int nchunks = (p.Lk16 + 31) / 32;
    for (int item = warp; item < mtl * nchunks; item += kWarps) {
      int nc, mt;
      split_item(item, mtl, nc, mt);
      float acc[4][4];
      // the last chunk may run 16 columns past Lk16: those B rows belong to the V tile (finite data), results are dropped
      warp_mma_16x32<false>(acc, Qs, hp, mt * 16, Ks, hp, nc * 32, p.hd, lane);
#pragma unroll
      for (int jn = 0; jn < 4; ++jn)
#pragma unroll
        for (int half = 0; half < 2; ++half) {
          const int m = mt * 16 + (lane >> 2) + 8 * half, n = nc * 32 + jn * 8 + (lane & 3) * 2;
          if (n < p.Lk16) *reinterpret_cast<float2*>(Sf + m * sp + n) = make_float2(acc[jn][2 * half] * p.scale, acc[jn][2 * half + 1] * p.scale);
        }
    }
    __syncthreads();
    if (qb + (int)gridDim.x < nqb) {      // the Q tile is free: fetch the next block's while this one finishes
      const int qn = (qb + gridDim.x) * kQB;
      load_rows(Qs, hp, qsrc, p.ldq, qn, p.Lq, kQB, p.hd, bar1);
    }

    // row softmax (one warp per row; lane l owns the column pairs 2l + 64 i), LSE out, dropout (one hash per pair), bf16 Pd
    // written over the row's own fp32 storage
    for (int r = warp; r < mtl * 16; r += kWarps) {
      float2 v[kMaxLk / 64];
      float mx = -INFINITY;
      float* srow = Sf + r * sp;
#pragma unroll
      for (int i = 0; i < kMaxLk / 64; ++i) {
        const int jx = 2 * lane + 64 * i;
        v[i] = (jx < p.Lk16) ? *reinterpret_cast<const float2*>(srow + jx) : make_float2(-INFINITY, -INFINITY);
        if (jx >= p.Lk) v[i].x = -INFINITY;
        if (jx + 1 >= p.Lk) v[i].y = -INFINITY;
        mx = fmaxf(mx, fmaxf(v[i].x, v[i].y));
      }
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, o));
      float sum = 0.f;
      const float mxl = mx * 1.4426950408889634f;
#pragma unroll
      for (int i = 0; i < kMaxLk / 64; ++i) {
        v[i].x = psg_ex2_approx(fmaf(v[i].x, 1.4426950408889634f, -mxl));       // exp(-inf) = 0 on the padding
        v[i].y = psg_ex2_approx(fmaf(v[i].y, 1.4426950408889634f, -mxl));
        sum += v[i].x + v[i].y;
      }
      sum = psg_warp_sum(sum);
      const float inv = 1.f / sum;
      const bool valid = q0 + r < p.Lq;
      const long long row_global = ((long long)bh) * p.Lq + q0 + r;
      if (valid && lane == 0 && p.lse) p.lse[row_global] = mx + __logf(sum);
      __syncwarp();
      __nv_bfloat16* prow = reinterpret_cast<__nv_bfloat16*>(srow);
      const uint64_t row_base = (uint64_t)row_global * (uint64_t)p.Lk;
      const float live = valid ? inv : 0.f;
#pragma unroll
      for (int i = 0; i < kMaxLk / 64; ++i) {
        const int jx = 2 * lane + 64 * i;
        if (jx < p.Lk16) {
          float p0 = v[i].x * live, p1 = v[i].y * live;
          if (p.thr) {
            bool k0, k1;
            keep_pair(p, row_base, jx, k0, k1);
            p0 = k0 ? p0 * p.ks : 0.f;
            p1 = k1 ? p1 * p.ks : 0.f;
          }
          *reinterpret_cast<__nv_bfloat162*>(prow + jx) = __floats2bfloat162_rn(p0, p1);
        }
      }
    }
    __syncthreads();

    // O = Pd V
    const __nv_bfloat16* Pd = reinterpret_cast<const __nv_bfloat16*>(Sf);
    const int ldp = 2 * sp;
    __nv_bfloat16* obase = p.out + (long long)b * p.Lq * p.ldo + h * p.hd;
    const int ochunks = (p.hd + 31) / 32;      // head_dim % 32 == 16: the last chunk's upper half is dropped
    for (int item = warp; item < mtl * ochunks; item += kWarps) {
      int nc, mt;
      split_item(item, mtl, nc, mt);
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
}

// ---------------------------------------------------------------------------------------------------------------------
// backward, dQ
//   smem: K [Lk16][hp] | V [Lk16][hp] | Q [64][hp] | dO [64][hp] | dS bf16 [64][dp] | lse[64] | delta[64]
// ---------------------------------------------------------------------------------------------------------------------
constexpr int kBB = 64;

__global__ void __launch_bounds__(kThreads) attn_bwd_dq_kernel(const Params p) {
  extern __shared__ __align__(16) uint8_t smem[];
  const int bh = blockIdx.y, b = bh / p.H, h = bh - b * p.H;
  const int hp = p.hp, dp = p.Lk16 + 8;
  __nv_bfloat16* Ks = reinterpret_cast<__nv_bfloat16*>(smem);
  __nv_bfloat16* Vs = Ks + (size_t)p.Lk16 * hp;
  __nv_bfloat16* Qs = Vs + (size_t)p.Lk16 * hp;
  __nv_bfloat16* dOs = Qs + (size_t)kBB * hp;
  __nv_bfloat16* dSs = dOs + (size_t)kBB * hp;
  float* lse_s = reinterpret_cast<float*>(dSs + (size_t)kBB * dp);
  float* del_s = lse_s + kBB;
  const uint32_t bar0 = smem_u32(smem + p.bar_off), bar1 = bar0 + 8;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, kWarps = blockDim.x >> 5;
  const int nqb = (p.Lq + kBB - 1) / kBB;
  const __nv_bfloat16* qsrc = p.q + (long long)b * p.Lq * p.ldq + h * p.hd;
  const __nv_bfloat16* dosrc = p.dout + (long long)b * p.Lq * p.lddo + h * p.hd;
  const float scale_l2 = p.scale * 1.4426950408889634f;

  if (threadIdx.x == 0) {
    mbar_init(bar0, 1);
    mbar_init(bar1, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    mbar_expect_tx(bar0, 2u * tile_bytes(0, p.Lk, p.Lk16, p.hd));
    mbar_expect_tx(bar1, 2u * tile_bytes(blockIdx.x * kBB, p.Lq, kBB, p.hd));
  }
  __syncthreads();           // barriers initialised and their byte counts posted before any copy is issued
  load_rows(Ks, hp, p.k + (long long)b * p.Lk * p.ldk + h * p.hd, p.ldk, 0, p.Lk, p.Lk16, p.hd, bar0);
  load_rows(Qs, hp, qsrc, p.ldq, blockIdx.x * kBB, p.Lq, kBB, p.hd, bar1);
  load_rows(Vs, hp, p.v + (long long)b * p.Lk * p.ldv + h * p.hd, p.ldv, 0, p.Lk, p.Lk16, p.hd, bar0);
  load_rows(dOs, hp, dosrc, p.lddo, blockIdx.x * kBB, p.Lq, kBB, p.hd, bar1);

  uint32_t phase = 0;
  for (int qb = blockIdx.x; qb < nqb; qb += gridDim.x) {
    const int q0 = qb * kBB;
    const int mtl = min(kBB / 16, (p.Lq - q0 + 15) >> 4);      // live 16-row tiles of this block
    // delta[i] = dO_i . O_i (both straight from global, while the tiles are in flight), one warp per row; also stage lse
    for (int r = warp; r < kBB; r += kWarps) {
      const int qi = q0 + r;
      float s = 0.f;
      if (qi < p.Lq) {
        const __nv_bfloat16* orow = p.o + ((long long)b * p.Lq + qi) * p.ldo + h * p.hd;
        const __nv_bfloat16* drow = dosrc + (long long)qi * p.lddo;
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
        lse_s[r] = (qi < p.Lq) ? p.lse[rg] * 1.4426950408889634f : 0.f;
        if (qi < p.Lq) p.delta[rg] = s;
      }
    }
    if (qb == (int)blockIdx.x) mbar_wait(bar0, 0);
    mbar_wait(bar1, phase);
    phase ^= 1;
    if (threadIdx.x == 0 && qb + (int)gridDim.x < nqb) mbar_expect_tx(bar1, 2u * tile_bytes((qb + gridDim.x) * kBB, p.Lq, kBB, p.hd));
    __syncthreads();

    // per 16 x 32 tile: S = Q K^T and dPd = dO V^T in registers -> dS = P o (drop(dPd) - delta) -> shared memory
    const int nchunks = (p.Lk16 + 31) / 32;
    for (int item = warp; item < mtl * nchunks; item += kWarps) {
      int nc, mt;
      split_item(item, mtl, nc, mt);
      float sa[4][4], da[4][4];
      warp_mma_16x32<false>(sa, Qs, hp, mt * 16, Ks, hp, nc * 32, p.hd, lane);
      warp_mma_16x32<false>(da, dOs, hp, mt * 16, Vs, hp, nc * 32, p.hd, lane);
#pragma unroll
      for (int half = 0; half < 2; ++half) {
        const int m = mt * 16 + (lane >> 2) + 8 * half;
        const bool row_ok = q0 + m < p.Lq;
        const float lse_m = lse_s[m], del_m = del_s[m];
        const uint64_t row_base = (uint64_t)((long long)bh * p.Lq + q0 + m) * (uint64_t)p.Lk;
        __nv_bfloat16* drow = dSs + m * dp;
#pragma unroll
        for (int jn = 0; jn < 4; ++jn) {
          const int n = nc * 32 + jn * 8 + (lane & 3) * 2;          // columns n, n + 1
          if (n >= p.Lk16) continue;
          float d0 = 0.f, d1 = 0.f;
          if (row_ok) {
            bool k0 = true, k1 = true;
            if (p.thr) keep_pair(p, row_base, n, k0, k1);
            if (n < p.Lk) d0 = psg_ex2_approx(fmaf(sa[jn][2 * half], scale_l2, -lse_m)) * ((k0 ? da[jn][2 * half] * p.ks : 0.f) - del_m);
            if (n + 1 < p.Lk) d1 = psg_ex2_approx(fmaf(sa[jn][2 * half + 1], scale_l2, -lse_m)) * ((k1 ? da[jn][2 * half + 1] * p.ks : 0.f) - del_m);
          }
          *reinterpret_cast<__nv_bfloat162*>(drow + n) = __floats2bfloat162_rn(d0, d1);
        }
      }
    }
    __syncthreads();
    if (qb + (int)gridDim.x < nqb) {      // Q / dO tiles are free: fetch the next block's under the dQ product
      const int qn = (qb + gridDim.x) * kBB;
      load_rows(Qs, hp, qsrc, p.ldq, qn, p.Lq, kBB, p.hd, bar1);
      load_rows(dOs, hp, dosrc, p.lddo, qn, p.Lq, kBB, p.hd, bar1);
    }
    // dQ = scale * dS K
    __nv_bfloat16* qbase = p.dq + (long long)b * p.Lq * p.lddq + h * p.hd;
    const int ochunks = (p.hd + 31) / 32;      // head_dim % 32 == 16: the last chunk's upper half is dropped
    for (int item = warp; item < mtl * ochunks; item += kWarps) {
      int nc, mt;
      split_item(item, mtl, nc, mt);
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
    __syncthreads();           // lse_s / del_s / dSs are rewritten by the next block
  }
}

// ---------------------------------------------------------------------------------------------------------------------
// backward, dK / dV
//   smem: Q [Lq16][hp] | dO [Lq16][hp] | K [64][hp] | V [64][hp] | T bf16 [64][dp] (Pd^T, then dS^T) | lse[Lq16] | delta[Lq16]
// ---------------------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(kThreads) attn_bwd_dkv_kernel(const Params p) {
  extern __shared__ __align__(16) uint8_t smem[];
  const int bh = blockIdx.y, b = bh / p.H, h = bh - b * p.H;
  const int hp = p.hp, dp = p.Lq16 + 8;
  __nv_bfloat16* Qs = reinterpret_cast<__nv_bfloat16*>(smem);
  __nv_bfloat16* dOs = Qs + (size_t)p.Lq16 * hp;
  __nv_bfloat16* Ks = dOs + (size_t)p.Lq16 * hp;
  __nv_bfloat16* Vs = Ks + (size_t)kBB * hp;
  __nv_bfloat16* Ts = Vs + (size_t)kBB * hp;
  float* lse_s = reinterpret_cast<float*>(Ts + (size_t)kBB * dp);
  float* del_s = lse_s + p.Lq16;
  const uint32_t bar0 = smem_u32(smem + p.bar_off), bar1 = bar0 + 8;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, kWarps = blockDim.x >> 5;
  const int nkb = (p.Lk + kBB - 1) / kBB;
  const __nv_bfloat16* ksrc = p.k + (long long)b * p.Lk * p.ldk + h * p.hd;
  const __nv_bfloat16* vsrc = p.v + (long long)b * p.Lk * p.ldv + h * p.hd;
  const float scale_l2 = p.scale * 1.4426950408889634f;

  if (threadIdx.x == 0) {
    mbar_init(bar0, 1);
    mbar_init(bar1, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    mbar_expect_tx(bar0, 2u * tile_bytes(0, p.Lq, p.Lq16, p.hd));
    mbar_expect_tx(bar1, 2u * tile_bytes(blockIdx.x * kBB, p.Lk, kBB, p.hd));
  }
  __syncthreads();           // barriers initialised and their byte counts posted before any copy is issued
  load_rows(Ks, hp, ksrc, p.ldk, blockIdx.x * kBB, p.Lk, kBB, p.hd, bar1);
  load_rows(Qs, hp, p.q + (long long)b * p.Lq * p.ldq + h * p.hd, p.ldq, 0, p.Lq, p.Lq16, p.hd, bar0);
  load_rows(Vs, hp, vsrc, p.ldv, blockIdx.x * kBB, p.Lk, kBB, p.hd, bar1);
  load_rows(dOs, hp, p.dout + (long long)b * p.Lq * p.lddo + h * p.hd, p.lddo, 0, p.Lq, p.Lq16, p.hd, bar0);
  for (int i = threadIdx.x; i < p.Lq16; i += blockDim.x) {
    const bool ok = i < p.Lq;
    lse_s[i] = ok ? p.lse[(long long)bh * p.Lq + i] * 1.4426950408889634f : 0.f;
    del_s[i] = ok ? p.delta[(long long)bh * p.Lq + i] : 0.f;
  }
  mbar_wait(bar0, 0);

  const int nchunks = (p.Lq16 + 31) / 32;
  const int ochunks = (p.hd + 31) / 32;
  uint32_t phase = 0;
  for (int kb = blockIdx.x; kb < nkb; kb += gridDim.x) {
    const int k0 = kb * kBB;
    const int mtl = min(kBB / 16, (p.Lk - k0 + 15) >> 4);      // live 16-key tiles of this block
    mbar_wait(bar1, phase);
    phase ^= 1;
    if (threadIdx.x == 0 && kb + (int)gridDim.x < nkb) mbar_expect_tx(bar1, 2u * tile_bytes((kb + gridDim.x) * kBB, p.Lk, kBB, p.hd));
    __syncthreads();
    // pass A: Pd^T = drop(exp(scale K Q^T - lse)) -> T (rows = keys of this block, columns = queries)
    for (int item = warp; item < mtl * nchunks; item += kWarps) {
      int nc, mt;
      split_item(item, mtl, nc, mt);
      float sa[4][4];
      warp_mma_16x32<false>(sa, Ks, hp, mt * 16, Qs, hp, nc * 32, p.hd, lane);
#pragma unroll
      for (int jn = 0; jn < 4; ++jn)
#pragma unroll
        for (int half = 0; half < 2; ++half) {
          const int m = mt * 16 + (lane >> 2) + 8 * half, n = nc * 32 + jn * 8 + (lane & 3) * 2;    // queries n, n + 1
          if (n >= p.Lq16) continue;
          float pd0 = 0.f, pd1 = 0.f;
          if (k0 + m < p.Lk) {
            if (n < p.Lq) {
              pd0 = psg_ex2_approx(fmaf(sa[jn][2 * half], scale_l2, -lse_s[n]));
              if (p.thr) pd0 = keep_ij(p, (long long)bh * p.Lq + n, k0 + m) ? pd0 * p.ks : 0.f;
            }
            if (n + 1 < p.Lq) {
              pd1 = psg_ex2_approx(fmaf(sa[jn][2 * half + 1], scale_l2, -lse_s[n + 1]));
              if (p.thr) pd1 = keep_ij(p, (long long)bh * p.Lq + n + 1, k0 + m) ? pd1 * p.ks : 0.f;
            }
          }
          *reinterpret_cast<__nv_bfloat162*>(Ts + m * dp + n) = __floats2bfloat162_rn(pd0, pd1);
        }
    }
    __syncthreads();
    // dV = Pd^T dO
    __nv_bfloat16* vbase = p.dv + (long long)b * p.Lk * p.lddv + h * p.hd;
    for (int item = warp; item < mtl * ochunks; item += kWarps) {
      int nc, mt;
      split_item(item, mtl, nc, mt);
      float acc[4][4];
      warp_mma_16x32<true>(acc, Ts, dp, mt * 16, dOs, hp, nc * 32, p.Lq16, lane);
#pragma unroll
      for (int jn = 0; jn < 4; ++jn)
#pragma unroll
        for (int half = 0; half < 2; ++half) {
          const int m = k0 + mt * 16 + (lane >> 2) + 8 * half, n = nc * 32 + jn * 8 + (lane & 3) * 2;
          if (m < p.Lk && n < p.hd)
            *reinterpret_cast<__nv_bfloat162*>(vbase + (long long)m * p.lddv + n) = __floats2bfloat162_rn(acc[jn][2 * half], acc[jn][2 * half + 1]);
        }
    }
    __syncthreads();
    // pass B: dS^T = P^T o (drop(V dO^T) - delta) -> T
    for (int item = warp; item < mtl * nchunks; item += kWarps) {
      int nc, mt;
      split_item(item, mtl, nc, mt);
      float sa[4][4], da[4][4];
      warp_mma_16x32<false>(sa, Ks, hp, mt * 16, Qs, hp, nc * 32, p.hd, lane);
      warp_mma_16x32<false>(da, Vs, hp, mt * 16, dOs, hp, nc * 32, p.hd, lane);
#pragma unroll
      for (int jn = 0; jn < 4; ++jn)
#pragma unroll
        for (int half = 0; half < 2; ++half) {
          const int m = mt * 16 + (lane >> 2) + 8 * half, n = nc * 32 + jn * 8 + (lane & 3) * 2;    // queries n, n + 1
          if (n >= p.Lq16) continue;
          float ds0 = 0.f, ds1 = 0.f;
          if (k0 + m < p.Lk) {
            if (n < p.Lq) {
              const float pr = psg_ex2_approx(fmaf(sa[jn][2 * half], scale_l2, -lse_s[n]));
              float dpv = da[jn][2 * half];
              if (p.thr) dpv = keep_ij(p, (long long)bh * p.Lq + n, k0 + m) ? dpv * p.ks : 0.f;
              ds0 = pr * (dpv - del_s[n]);
            }
            if (n + 1 < p.Lq) {
              const float pr = psg_ex2_approx(fmaf(sa[jn][2 * half + 1], scale_l2, -lse_s[n + 1]));
              float dpv = da[jn][2 * half + 1];
              if (p.thr) dpv = keep_ij(p, (long long)bh * p.Lq + n + 1, k0 + m) ? dpv * p.ks : 0.f;
              ds1 = pr * (dpv - del_s[n + 1]);
            }
          }
          *reinterpret_cast<__nv_bfloat162*>(Ts + m * dp + n) = __floats2bfloat162_rn(ds0, ds1);
        }
    }
    __syncthreads();
    if (kb + (int)gridDim.x < nkb) {      // K / V tiles are free: fetch the next block's under the dK product
      const int kn = (kb + gridDim.x) * kBB;
      load_rows(Ks, hp, ksrc, p.ldk, kn, p.Lk, kBB, p.hd, bar1);
      load_rows(Vs, hp, vsrc, p.ldv, kn, p.Lk, kBB, p.hd, bar1);
    }
    // dK = scale * dS^T Q
    __nv_bfloat16* kbase = p.dk + (long long)b * p.Lk * p.lddk + h * p.hd;
    for (int item = warp; item < mtl * ochunks; item += kWarps) {
      int nc, mt;
      split_item(item, mtl, nc, mt);
      float acc[4][4];
      warp_mma_16x32<true>(acc, Ts, dp, mt * 16, Qs, hp, nc * 32, p.Lq16, lane);
#pragma unroll
      for (int jn = 0; jn < 4; ++jn)
#pragma unroll
        for (int half = 0; half < 2; ++half) {
          const int m = k0 + mt * 16 + (lane >> 2) + 8 * half, n = nc * 32 + jn * 8 + (lane & 3) * 2;
          if (m < p.Lk && n < p.hd)
            *reinterpret_cast<__nv_bfloat162*>(kbase + (long long)m * p.lddk + n) =
                __floats2bfloat162_rn(acc[jn][2 * half] * p.scale, acc[jn][2 * half + 1] * p.scale);
        }
    }
    // (the next block's first __syncthreads orders these reads of T before its pass A writes)
  }
}

// ---------------------------------------------------------------------------------------------------------------------
// backward for small problems (Lq <= 64 and Lk <= 64: the 7x7 and 4x4 levels and their cross-attention): ONE CTA per
// (batch, head) holds Q, K, V, dO, computes S / dPd once and produces dQ, dK and dV (the two-kernel path loads the same four
// tiles twice, recomputes S three times and spends most of its time in launch and phase latency at these sizes).
//   smem: K [Lk16][hp] | V [Lk16][hp] | Q [Lq16][hp] | dO [Lq16][hp] | Pd bf16 [Lq16][dp] | dS bf16 [Lq16][dp] | lse | delta
//   dQ = scale dS K;  dV = Pd^T dO;  dK = scale dS^T Q   (the transposed products read Pd / dS through ldmatrix.trans)
// ---------------------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(kThreads) attn_bwd_small_kernel(const Params p) {
  extern __shared__ __align__(16) uint8_t smem[];
  const int bh = blockIdx.x, b = bh / p.H, h = bh - b * p.H;
  const int hp = p.hp, dp = p.Lk16 + 8;
  __nv_bfloat16* Ks = reinterpret_cast<__nv_bfloat16*>(smem);
  __nv_bfloat16* Vs = Ks + (size_t)p.Lk16 * hp;
  __nv_bfloat16* Qs = Vs + (size_t)p.Lk16 * hp;
  __nv_bfloat16* dOs = Qs + (size_t)p.Lq16 * hp;
  __nv_bfloat16* Pds = dOs + (size_t)p.Lq16 * hp;
  __nv_bfloat16* dSs = Pds + (size_t)p.Lq16 * dp;
  float* lse_s = reinterpret_cast<float*>(dSs + (size_t)p.Lq16 * dp);
  float* del_s = lse_s + p.Lq16;
  const uint32_t bar0 = smem_u32(smem + p.bar_off);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, kWarps = blockDim.x >> 5;
  const __nv_bfloat16* dosrc = p.dout + (long long)b * p.Lq * p.lddo + h * p.hd;
  const float scale_l2 = p.scale * 1.4426950408889634f;

  if (threadIdx.x == 0) {
    mbar_init(bar0, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    mbar_expect_tx(bar0, 2u * tile_bytes(0, p.Lk, p.Lk16, p.hd) + 2u * tile_bytes(0, p.Lq, p.Lq16, p.hd));
  }
  __syncthreads();
  load_rows(Ks, hp, p.k + (long long)b * p.Lk * p.ldk + h * p.hd, p.ldk, 0, p.Lk, p.Lk16, p.hd, bar0);
  load_rows(Qs, hp, p.q + (long long)b * p.Lq * p.ldq + h * p.hd, p.ldq, 0, p.Lq, p.Lq16, p.hd, bar0);
  load_rows(Vs, hp, p.v + (long long)b * p.Lk * p.ldv + h * p.hd, p.ldv, 0, p.Lk, p.Lk16, p.hd, bar0);
  load_rows(dOs, hp, dosrc, p.lddo, 0, p.Lq, p.Lq16, p.hd, bar0);
  // delta[i] = dO_i . O_i straight from global while the tiles are in flight, one warp per row; also stage lse (log2 units)
  for (int r = warp; r < p.Lq16; r += kWarps) {
    float s = 0.f;
    if (r < p.Lq) {
      const __nv_bfloat16* orow = p.o + ((long long)b * p.Lq + r) * p.ldo + h * p.hd;
      const __nv_bfloat16* drow = dosrc + (long long)r * p.lddo;
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
      del_s[r] = s;
      lse_s[r] = (r < p.Lq) ? p.lse[(long long)bh * p.Lq + r] * 1.4426950408889634f : 0.f;
    }
  }
  mbar_wait(bar0, 0);
  __syncthreads();

  // S = Q K^T and dPd = dO V^T per 16 x 32 tile in registers -> Pd = drop(P), dS = P o (drop(dPd) - delta) -> shared memory
  const int mq = p.Lq16 >> 4, mk = p.Lk16 >> 4;
  const int nchunks = (p.Lk16 + 31) / 32;
  for (int item = warp; item < mq * nchunks; item += kWarps) {
    const int nc = item / mq, mt = item - nc * mq;
    float sa[4][4], da[4][4];
    // (a chunk may run 16 columns past Lk16: those B rows belong to the next tile in shared memory, results are dropped)
    warp_mma_16x32<false>(sa, Qs, hp, mt * 16, Ks, hp, nc * 32, p.hd, lane);
    warp_mma_16x32<false>(da, dOs, hp, mt * 16, Vs, hp, nc * 32, p.hd, lane);
#pragma unroll
    for (int half = 0; half < 2; ++half) {
      const int m = mt * 16 + (lane >> 2) + 8 * half;
      const bool row_ok = m < p.Lq;
      const float lse_m = lse_s[m], del_m = del_s[m];
      const uint64_t row_base = (uint64_t)((long long)bh * p.Lq + m) * (uint64_t)p.Lk;
#pragma unroll
      for (int jn = 0; jn < 4; ++jn) {
        const int n = nc * 32 + jn * 8 + (lane & 3) * 2;          // columns n, n + 1
        if (n >= p.Lk16) continue;
        float p0 = 0.f, p1 = 0.f, d0 = 0.f, d1 = 0.f;
        if (row_ok) {
          bool k0 = true, k1 = true;
          if (p.thr) keep_pair(p, row_base, n, k0, k1);
          if (n < p.Lk) {
            const float pr = psg_ex2_approx(fmaf(sa[jn][2 * half], scale_l2, -lse_m));
            p0 = k0 ? pr * p.ks : 0.f;
            d0 = pr * ((k0 ? da[jn][2 * half] * p.ks : 0.f) - del_m);
          }
          if (n + 1 < p.Lk) {
            const float pr = psg_ex2_approx(fmaf(sa[jn][2 * half + 1], scale_l2, -lse_m));
            p1 = k1 ? pr * p.ks : 0.f;
            d1 = pr * ((k1 ? da[jn][2 * half + 1] * p.ks : 0.f) - del_m);
          }
        }
        *reinterpret_cast<__nv_bfloat162*>(Pds + m * dp + n) = __floats2bfloat162_rn(p0, p1);
        *reinterpret_cast<__nv_bfloat162*>(dSs + m * dp + n) = __floats2bfloat162_rn(d0, d1);
      }
    }
  }
  __syncthreads();
  // the three output products, one item list: [dQ tiles | dV tiles | dK tiles] x head_dim chunks of 32
  const int ochunks = (p.hd + 31) / 32;      // head_dim % 32 == 16: the last chunk's upper half is dropped
  const int nq = mq * ochunks, nk = mk * ochunks;
  for (int item = warp; item < nq + 2 * nk; item += kWarps) {
    float acc[4][4];
    __nv_bfloat16* obase;
    long long ldo_;
    int mt, nc, rows;
    float osc;
    if (item < nq) {                     // dQ = scale dS K
      nc = item / mq; mt = item - nc * mq;
      warp_mma_16x32<true>(acc, dSs, dp, mt * 16, Ks, hp, nc * 32, p.Lk16, lane);
      obase = p.dq + (long long)b * p.Lq * p.lddq + h * p.hd; ldo_ = p.lddq; rows = p.Lq; osc = p.scale;
    } else if (item < nq + nk) {         // dV = Pd^T dO
      const int it = item - nq;
      nc = it / mk; mt = it - nc * mk;
      warp_mma_16x32_at(acc, Pds, dp, mt * 16, dOs, hp, nc * 32, p.Lq16, lane);
      obase = p.dv + (long long)b * p.Lk * p.lddv + h * p.hd; ldo_ = p.lddv; rows = p.Lk; osc = 1.f;
    } else {                             // dK = scale dS^T Q
      const int it = item - nq - nk;
      nc = it / mk; mt = it - nc * mk;
      warp_mma_16x32_at(acc, dSs, dp, mt * 16, Qs, hp, nc * 32, p.Lq16, lane);
      obase = p.dk + (long long)b * p.Lk * p.lddk + h * p.hd; ldo_ = p.lddk; rows = p.Lk; osc = p.scale;
    }
#pragma unroll
    for (int jn = 0; jn < 4; ++jn)
#pragma unroll
      for (int half = 0; half < 2; ++half) {
        const int m = mt * 16 + (lane >> 2) + 8 * half, n = nc * 32 + jn * 8 + (lane & 3) * 2;
        if (m < rows && n < p.hd)
          *reinterpret_cast<__nv_bfloat162*>(obase + (long long)m * ldo_ + n) = __floats2bfloat162_rn(acc[jn][2 * half] * osc, acc[jn][2 * half + 1] * osc);
      }
  }
}

static size_t small_smem(const Params& p) {
  return ((size_t)2 * p.Lk16 * p.hp + (size_t)2 * p.Lq16 * p.hp) * 2 + (size_t)2 * p.Lq16 * (p.Lk16 + 8) * 2 + (size_t)2 * p.Lq16 * 4 + 64;
}

static size_t fwd_smem(const Params& p) {
  return ((size_t)2 * p.Lk16 * p.hp + (size_t)kQB * p.hp) * 2 + (size_t)kQB * (p.Lk16 + 4) * 4 + 64;
}
static size_t dq_smem(const Params& p) {
  return ((size_t)2 * p.Lk16 * p.hp + (size_t)2 * kBB * p.hp) * 2 + (size_t)kBB * (p.Lk16 + 8) * 2 + 2 * kBB * 4 + 64;
}
static size_t dkv_smem(const Params& p) {
  return ((size_t)2 * p.Lq16 * p.hp + (size_t)2 * kBB * p.hp) * 2 + (size_t)kBB * (p.Lq16 + 8) * 2 + (size_t)2 * p.Lq16 * 4 + 64;
}

// one CTA per SM anyway (shared memory): 16 warps; otherwise 8 warps so that two or three CTAs share an SM
static int threads_for(size_t smem) { return smem > 110 * 1024 ? kThreads : 256; }

// CTAs per (batch, head): 1 when the heads alone fill the GPU (resident tiles fetched once per head), else as many as
// keep every SM busy.  g_split > 0 forces a value (psg_attn_fused_split: tests drive the block loop at small batch).
static int g_split = 0;
static int g_small_bwd = 1;      // psg_attn_fused_small_bwd: single-kernel backward for Lq, Lk <= 64
static int split_for(int nblocks, int heads_total) {
  int s = g_split > 0 ? g_split : (psg_num_sms() + heads_total - 1) / heads_total;
  return s < 1 ? 1 : (s > nblocks ? nblocks : s);
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

// attention_umma.cu: the tcgen05 / TMEM kernels that take the problems with >= 65 queries (the 196-token level)
int psg_attn_umma_ok(int B, int H, int Lq, int Lk, int hd);
int psg_attn_umma_fwd(const void* q, long long ldq, const void* k, long long ldk, const void* v, long long ldv, void* o, long long ldo,
                      float* lse, int B, int H, int Lq, int Lk, int hd, float scale, unsigned long long drop_seed, float drop_p,
                      void* stream);
int psg_attn_umma_bwd(const void* q, long long ldq, const void* k, long long ldk, const void* v, long long ldv, const void* o,
                      long long ldo, const void* dout, long long lddo, const float* lse, float* delta, void* dq, long long lddq,
                      void* dk, long long lddk, void* dv, long long lddv, int B, int H, int Lq, int Lk, int hd, float scale,
                      unsigned long long drop_seed, float drop_p, void* stream);

// Test / measurement hook: CTAs per (batch, head) of the fused attention kernels (0 = by problem size).  Returns the previous value.
int psg_attn_fused_split(int n) {
  const int prev = fattn::g_split;
  if (n >= 0) fattn::g_split = n;
  return prev;
}

// Test / measurement hook: 1 (default) = problems with Lq, Lk <= 64 use the single-kernel backward; 0 = always the dQ + dK/dV
// pair.  Returns the previous value.
int psg_attn_fused_small_bwd(int on) {
  const int prev = fattn::g_small_bwd;
  if (on == 0 || on == 1) fattn::g_small_bwd = on;
  return prev;
}

// 1 if the fused kernels take this problem (bf16; head_dim % 16 == 0; Lk <= 256; everything fits in shared memory).
int psg_attn_fused_ok(int B, int H, int Lq, int Lk, int hd) {
  if (psg_attn_umma_ok(B, H, Lq, Lk, hd)) return 1;
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
  if (psg_attn_umma_ok(B, H, Lq, Lk, hd) && ldo % 8 == 0 && ((uintptr_t)o % 16 == 0))
    return psg_attn_umma_fwd(q, ldq, k, ldk, v, ldv, o, ldo, lse, B, H, Lq, Lk, hd, scale, drop_seed, drop_p, stream);
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
  dim3 grid(split_for((Lq + kQB - 1) / kQB, B * H), B * H);
  p.bar_off = (int)fwd_smem(p) - 32;
  attn_fwd_kernel<<<grid, threads_for(fwd_smem(p)), fwd_smem(p), (cudaStream_t)stream>>>(p);
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
  if (psg_attn_umma_ok(B, H, Lq, Lk, hd) && lddq % 8 == 0 && lddk % 8 == 0 && lddv % 8 == 0 && ((uintptr_t)dq % 16 == 0) &&
      ((uintptr_t)dk % 16 == 0) && ((uintptr_t)dv % 16 == 0))
    return psg_attn_umma_bwd(q, ldq, k, ldk, v, ldv, o, ldo, dout, lddo, lse, delta, dq, lddq, dk, lddk, dv, lddv, B, H, Lq, Lk, hd, scale,
                             drop_seed, drop_p, stream);
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
  cudaStream_t st0 = (cudaStream_t)stream;
  if (g_small_bwd && Lq <= 64 && Lk <= 64 && small_smem(p) <= kSmemLimit) {
    static bool done0 = false;
    int rc0 = configure(attn_bwd_small_kernel, done0, "psg_attn_fused_bwd");
    if (rc0) return rc0;
    p.bar_off = (int)small_smem(p) - 32;
    attn_bwd_small_kernel<<<B * H, threads_for(small_smem(p)), small_smem(p), st0>>>(p);
    PSG_CHECK_LAUNCH("psg_attn_fused_bwd");
    return PSG_OK;
  }
  static bool done1 = false, done2 = false;
  int rc = configure(attn_bwd_dq_kernel, done1, "psg_attn_fused_bwd");
  if (rc) return rc;
  rc = configure(attn_bwd_dkv_kernel, done2, "psg_attn_fused_bwd");
  if (rc) return rc;
  cudaStream_t st = (cudaStream_t)stream;
  p.bar_off = (int)dq_smem(p) - 32;
  attn_bwd_dq_kernel<<<dim3(split_for((Lq + kBB - 1) / kBB, B * H), B * H), threads_for(dq_smem(p)), dq_smem(p), st>>>(p);
  p.bar_off = (int)dkv_smem(p) - 32;
  attn_bwd_dkv_kernel<<<dim3(split_for((Lk + kBB - 1) / kBB, B * H), B * H), threads_for(dkv_smem(p)), dkv_smem(p), st>>>(p);
  PSG_CHECK_LAUNCH("psg_attn_fused_bwd");
  g_psg_launch_count += 1;  // two kernels
  return PSG_OK;
}

}  // extern "C"
