// tcgen05 / TMEM / TMA attention for sm_100a (bf16 in, fp32 accumulate): softmax(scale Q K^T) [dropout] V, and its backward.
// SURVEY.md §2.1 K6/K7; replaces the nn.MultiheadAttention core (src/models/unet.py:160-173,217,232-239: scale 1/sqrt(head_dim),
// softmax over keys, dropout on the probabilities) for the levels whose sequences fill a 128-row MMA tile (196 tokens).
//
// Every product runs on the 5th-generation tensor cores; scores / probabilities / their gradients never leave the SM:
//
//   forward   per (batch, head), per 128-query tile:
//               S = Q K^T                 tcgen05.mma  A = Q tile (smem, K-major)   B = K (smem, K-major)    D = TMEM fp32
//               P = exp2(..) [dropout]    softmax warps: tcgen05.ld S -> registers -> bf16 -> tcgen05.st P  (TMEM)
//               O = P V                   tcgen05.mma  A = P (TMEM)                 B = V (smem, MN-major)   D = TMEM fp32
//               O *= ks / rowsum          epilogue warps: tcgen05.ld -> bf16 -> global
//   dQ        per 128-query tile:  S = Q K^T, dP = dO V^T -> dS = P o (drop(dP) - delta) (bf16, written over S in TMEM)
//               dQ = scale dS K           A = dS (TMEM)   B = K (the same smem bytes, read MN-major)
//   dK / dV   per 128-key tile, per block of <= 96 queries:  S^T = K Q^T, dP^T = V dO^T -> Pd^T, dS^T (bf16, in place)
//               dV += Pd^T dO, dK += dS^T Q    A = TMEM   B = dO / Q (the resident smem tiles, read MN-major)
//
// Operand tiles are fetched by TMA from the packed projection outputs through 4-D tensor maps (head_dim, head, token, batch):
// a box is (W head_dim columns) x (rows of one sample), so rows past the end of a sample and columns past head_dim arrive as
// zeros.  W = 64 (SWIZZLE_128B) when head_dim % 64 == 0, else 32 (SWIZZLE_64B: head_dim 160 = 5 boxes, no padding).  A box
// [rows][W] is at the same time a K-major operand (rows = M/N index) and an MN-major operand (rows = K index): K, Q and dO
// are used both ways without a second copy.
//
// Warp roles (320 threads, one CTA per SM, persistent over (batch, head) units): warp 0 TMA producer, warp 1 MMA issuer,
// warps 2..9 the TMEM <-> register work (a warp may only touch the 32 TMEM lanes of its quarter, warp % 4).
//
// The dropout mask is the library-wide stateless rule psg_drop_keep(seed, ((b*H + h)*Lq + i)*Lk + j), so these kernels,
// the mma.sync kernels of attention_fused.cu and the unfused fallback drop the same elements.
#include "psg_common.cuh"
#include <cuda.h>
#include <string.h>

namespace uattn {

constexpr int kThreads = 448;          // warp 0 TMA, warp 1 MMA, warps 2..9 TMEM<->register work, warps 10..13 epilogue / statistics
constexpr int kTileM = 128;
constexpr int kQBlk = 96;            // dK/dV kernel: queries per block (TMEM: 2 * 96 score columns + 2 * head_dim accumulators)
constexpr size_t kSmemLimit = 232448;

__device__ int g_timeout_flag = 0;

// Per-role cycle accounting for tuning (compiled in with -DUATTN_PROF only): lane 0 of one warp per role accumulates the cycles
// between laps; psg_attn_umma_prof() reads the per-CTA totals back.
#ifdef UATTN_PROF
__device__ long long g_prof[160 * 32];
struct Prof {
  long long t, acc[8];
  __device__ __forceinline__ static long long now() { long long c; asm volatile("mov.u64 %0, %%clock64;" : "=l"(c)); return c; }
  __device__ __forceinline__ Prof() : t(now()) { for (int i = 0; i < 8; ++i) acc[i] = 0; }
  __device__ __forceinline__ void lap(int i) { const long long n = now(); acc[i] += n - t; t = n; }
  __device__ __forceinline__ void flush(int role) { for (int i = 0; i < 8; ++i) g_prof[blockIdx.x * 32 + role * 8 + i] = acc[i]; }
};
#define PROF_DECL Prof prof
#define PROF_LAP(i) prof.lap(i)
#define PROF_FLUSH(role, cond) if (cond) prof.flush(role)
#else
#define PROF_DECL
#define PROF_LAP(i)
#define PROF_FLUSH(role, cond)
#endif

// ---------------------------------------------------------------------------------------------------------------------
// PTX wrappers
// ---------------------------------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
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
// Bounded waits: a protocol bug must never hang the GPU; on timeout raise a flag the host reads (psg_attn_umma_timeout_flag).
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
  for (uint32_t i = 0; i < (1u << 24); ++i)
    if (mbar_try_wait(bar, parity)) return;
  atomicExch(&g_timeout_flag, 1);
}
__device__ __forceinline__ void mbar_wait_backoff(uint32_t bar, uint32_t parity) {
  for (uint32_t i = 0; i < (1u << 22); ++i) {
    if (mbar_try_wait(bar, parity)) return;
    __nanosleep(64);
  }
  atomicExch(&g_timeout_flag, 1);
}
__device__ __forceinline__ void tma_load_4d(uint32_t dst, const CUtensorMap* tm, uint32_t bar, int c0, int c1, int c2, int c3) {
  asm volatile(
      "cp.async.bulk.tensor.4d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2];" ::"r"(dst),
      "l"(reinterpret_cast<uint64_t>(tm)), "r"(bar), "r"(c0), "r"(c1), "r"(c2), "r"(c3)
      : "memory");
}
__device__ __forceinline__ void prefetch_tmap(const CUtensorMap* tm) {
  asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(tm)) : "memory");
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tmem_alloc(uint32_t dst_smem, uint32_t ncols) {
  asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(dst_smem), "r"(ncols) : "memory");
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr, uint32_t ncols) {
  asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}
// D[tmem] (+)= A[smem desc] * B[smem desc]
__device__ __forceinline__ void umma_ss(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accum) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}" ::"r"(tmem_d),
      "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accum)
      : "memory");
}
// D[tmem] (+)= A[tmem: lane = row, one 32-bit column = two consecutive bf16 K elements] * B[smem desc]
__device__ __forceinline__ void umma_ts(uint32_t tmem_d, uint32_t tmem_a, uint64_t bdesc, uint32_t idesc, uint32_t accum) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t}" ::"r"(tmem_d),
      "r"(tmem_a), "l"(bdesc), "r"(idesc), "r"(accum)
      : "memory");
}
__device__ __forceinline__ void umma_commit(uint32_t bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void tmem_wait_ld() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }
__device__ __forceinline__ void tmem_wait_st() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, uint32_t (&r)[32]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]),
        "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]),
        "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
      : "r"(taddr)
      : "memory");
}
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, uint32_t (&r)[16]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
      : "r"(taddr)
      : "memory");
}
__device__ __forceinline__ void tmem_st16(uint32_t taddr, const uint32_t (&r)[16]) {
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], "
      "{%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16};" ::"r"(taddr),
      "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]), "r"(r[8]), "r"(r[9]), "r"(r[10]),
      "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15])
      : "memory");
}
__device__ __forceinline__ void tmem_st8(uint32_t taddr, const uint32_t (&r)[8]) {
  asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};" ::"r"(taddr), "r"(r[0]), "r"(r[1]),
               "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7])
               : "memory");
}
template <int N> __device__ __forceinline__ void tmem_ld(uint32_t taddr, uint32_t (&r)[N]);
template <> __device__ __forceinline__ void tmem_ld<32>(uint32_t taddr, uint32_t (&r)[32]) { tmem_ld32(taddr, r); }
template <> __device__ __forceinline__ void tmem_ld<16>(uint32_t taddr, uint32_t (&r)[16]) { tmem_ld16(taddr, r); }
template <int N> __device__ __forceinline__ void tmem_st(uint32_t taddr, const uint32_t (&r)[N]);
template <> __device__ __forceinline__ void tmem_st<16>(uint32_t taddr, const uint32_t (&r)[16]) { tmem_st16(taddr, r); }
template <> __device__ __forceinline__ void tmem_st<8>(uint32_t taddr, const uint32_t (&r)[8]) { tmem_st8(taddr, r); }

__device__ __forceinline__ uint32_t pack_bf16(float lo, float hi) {
  const __nv_bfloat162 h = __floats2bfloat162_rn(lo, hi);
  return *reinterpret_cast<const uint32_t*>(&h);
}

// Shared-memory matrix descriptor (cute/arch/mma_sm100_desc.hpp bit layout).  swz: 2 = SWIZZLE_128B, 4 = SWIZZLE_64B.
//  K-major : rows of `rowb` bytes, 8-row groups `sbo` apart (= 8 * rowb), LBO unused.
//  MN-major: K rows of `rowb` bytes (W consecutive MN elements), 8-row groups `sbo` apart, W-element MN groups `lbo` apart.
__device__ __forceinline__ uint64_t make_desc(uint32_t saddr, uint32_t lbo_bytes, uint32_t sbo_bytes, uint32_t swz) {
  uint64_t d = 0;
  d |= (uint64_t)((saddr & 0x3FFFF) >> 4);
  d |= (uint64_t)((lbo_bytes >> 4) & 0x3FFF) << 16;
  d |= (uint64_t)((sbo_bytes >> 4) & 0x3FFF) << 32;
  d |= (uint64_t)1 << 46;  // descriptor version (Blackwell)
  d |= (uint64_t)swz << 61;
  return d;
}
__device__ __forceinline__ uint32_t make_idesc(int n, int a_mn, int b_mn) {
  return (1u << 4)                      // D = f32
         | (1u << 7) | (1u << 10)       // A, B = bf16
         | ((uint32_t)a_mn << 15) | ((uint32_t)b_mn << 16) | ((uint32_t)(n >> 3) << 17) | ((uint32_t)(kTileM >> 4) << 24);
}

// Operand geometry shared by the three kernels (see the header comment).
struct Geo {
  int W;           // head_dim columns per TMA box: 64 (SWIZZLE_128B) or 32 (SWIZZLE_64B)
  int rowb;        // W * 2 bytes: one smem row of a box
  int nbox;        // ceil(hd / W)
  int kshift;      // log2 of the 16-element K steps per box row (W / 16): 2 or 1
  uint32_t swz;    // descriptor layout type
};

// K-major descriptor of the 16 head_dim columns of K-step `k` inside a [rows][W] box stack at `base` (boxes `box_bytes` apart)
__device__ __forceinline__ uint64_t kmajor_desc(const Geo& g, uint32_t base, uint32_t box_bytes, int k, uint32_t row0) {
  const int box = k >> g.kshift, kin = k - (box << g.kshift);
  return make_desc(base + box * box_bytes + row0 * g.rowb + kin * 32, 16, 8 * g.rowb, g.swz);
}
// MN-major descriptor of rows [row0, row0 + 16) (the K index) over all head_dim columns (the N index, boxes = MN groups)
__device__ __forceinline__ uint64_t mnmajor_desc(const Geo& g, uint32_t base, uint32_t box_bytes, uint32_t row0) {
  return make_desc(base + row0 * g.rowb, box_bytes, 8 * g.rowb, g.swz);
}

// MMA issue loops with the descriptors advanced incrementally.  Building both 64-bit descriptors from scratch cost the single issuing
// thread ~100 cycles per MMA -- more than a narrow MMA takes to execute -- and that issue time sat on the MMA -> elementwise -> MMA
// chain of every block (profiles/r02_attention_umma_role_cycles.txt: "mma: issue ...").  The start-address field is the low 14 bits
// (address >> 4) and shared memory is < 256 KB, so advancing a descriptor is one 64-bit add of (byte offset >> 4).
//   D (+)= A[128 rows, K-major boxes] x B[N rows from row_b, K-major boxes]^T over `ksteps` 16-wide K steps
__device__ __forceinline__ void mma_ss_loop(const Geo& g, uint32_t tmem_d, uint32_t base_a, uint32_t box_a, uint32_t base_b, uint32_t box_b,
                                            uint32_t row_b, int ksteps, uint32_t idesc) {
  uint64_t da = make_desc(base_a, 16, 8 * g.rowb, g.swz), db = make_desc(base_b + row_b * g.rowb, 16, 8 * g.rowb, g.swz);
  const int kpb = 1 << g.kshift;                       // K steps per box row
  const uint32_t jump_a = (box_a >> 4) - 2u * (kpb - 1), jump_b = (box_b >> 4) - 2u * (kpb - 1);
  int kin = 0;
  for (int k = 0; k < ksteps; ++k) {
    umma_ss(tmem_d, da, db, idesc, k > 0 ? 1u : 0u);
    if (++kin == kpb) { kin = 0; da += jump_a; db += jump_b; }
    else { da += 2; db += 2; }
  }
}
//   D (+)= A[TMEM, one K step every `a_step` columns] x B[rows row_b + 16 k of an MN-major box stack] over `ksteps` K steps
__device__ __forceinline__ void mma_ts_loop(const Geo& g, uint32_t tmem_d, uint32_t tmem_a, uint32_t a_step, uint32_t base_b, uint32_t box_b,
                                            uint32_t row_b, int ksteps, uint32_t idesc, bool accumulate_first) {
  uint64_t db = make_desc(base_b + row_b * g.rowb, box_b, 8 * g.rowb, g.swz);
  const uint32_t step_b = (16u * g.rowb) >> 4;
  for (int k = 0; k < ksteps; ++k) {
    umma_ts(tmem_d, tmem_a, db, idesc, (accumulate_first || k > 0) ? 1u : 0u);
    tmem_a += a_step;
    db += step_b;
  }
}

// The (unit, tile) walk every role performs identically: units blockIdx.x, + gridDim.x, ...; tiles 0 .. n_t - 1 of each.
struct Walk {
  int u, t, n_t, units, step;
  __device__ __forceinline__ Walk(int n_t_, int units_) : u(blockIdx.x), t(0), n_t(n_t_), units(units_), step(gridDim.x) {}
  __device__ __forceinline__ bool valid() const { return u < units; }
  __device__ __forceinline__ bool first() const { return t == 0; }
  __device__ __forceinline__ bool last() const { return t == n_t - 1; }
  __device__ __forceinline__ void advance() { if (++t == n_t) { t = 0; u += step; } }
};

// Dropout.  The library-wide rule is keep(idx) <=> 16-bit field (idx & 1) of psg_hash32(seed, idx >> 1) >= thr >> 16.  psg_hash32
// mixes the 64-bit pair index and the 64-bit seed into one 32-bit word first; the seed part is folded on the host (seedmix) and,
// when every pair index of the problem fits 32 bits (idx32: B*H*Lq*Lk < 2^33, always true for this model), so is the high word:
// the per-pair cost is one add, one xor and the two multiply / xorshift rounds.
struct Drop {
  unsigned long long seed;
  unsigned int thr;        // 0 = no dropout
  unsigned int seedmix;    // (uint32)seed ^ (uint32)(seed >> 32) * 0x85EBCA6B
  int idx32;
};
__device__ __forceinline__ uint32_t mix32(uint32_t x) {
  x ^= x >> 16;
  x *= 0x7FEB352Du;
  x ^= x >> 15;
  x *= 0x846CA68Bu;
  x ^= x >> 16;
  return x;
}
// zero the dropped ones of N consecutive elements starting at flat index `base` (one hash per aligned pair)
template <int N>
__device__ __forceinline__ void drop_mask(float (&e)[N], const Drop& dr, uint64_t base) {
  const uint32_t thr16 = dr.thr >> 16;
  if ((base & 1) == 0 && dr.idx32) {
    const uint32_t p0 = (uint32_t)(base >> 1);
#pragma unroll
    for (int j = 0; j < N; j += 2) {
      const uint32_t h = mix32((p0 + (uint32_t)(j >> 1)) ^ dr.seedmix);
      if ((h & 0xFFFFu) < thr16) e[j] = 0.f;
      if ((h >> 16) < thr16) e[j + 1] = 0.f;
    }
  } else if ((base & 1) == 0) {
#pragma unroll
    for (int j = 0; j < N; j += 2) {
      const uint32_t h = psg_hash32(dr.seed, (base + j) >> 1);
      if ((h & 0xFFFFu) < thr16) e[j] = 0.f;
      if ((h >> 16) < thr16) e[j + 1] = 0.f;
    }
  } else {
#pragma unroll
    for (int j = 0; j < N; ++j)
      if (!psg_drop_keep(dr.seed, base + j, dr.thr)) e[j] = 0.f;
  }
}

// =====================================================================================================================
// forward
//   warps 2..9: softmax (two warps per TMEM lane quarter: they interleave the 32-column chunks of the row and exchange their
//   partial row maxima through shared memory); warps 10..13: epilogue.
//   TMEM: S [0, Lk16) | P bf16 pairs [Lk16, 1.5 Lk16) | O [col_o, col_o + hd)
// =====================================================================================================================
struct FwdParams {
  CUtensorMap tm_q, tm_k, tm_v;
  __nv_bfloat16* out;
  long long ldo;
  float* lse;               // [B, H, Lq] or null
  int B, H, Lq, Lk, hd, Lk16;
  Geo g;
  int n_mt, units;
  float scale, scale_l2, ks;
  Drop drop;
  int col_p, col_o;
  uint32_t tmem_cols;
  int wide_out;             // output rows are 32-byte aligned: 256-bit stores
};

// F_R_FULL0/1: the row statistics of even / odd tiles are in shared memory.  Two barriers, because the epilogue may reach its
// wait a whole softmax pass late: with one barrier the phase it waits for could already be two behind (same parity).
enum { F_K_FULL, F_K_EMPTY, F_V_FULL, F_V_EMPTY, F_Q_FULL, F_Q_EMPTY, F_S_FULL, F_P_FULL, F_P_EMPTY, F_O_FULL, F_O_EMPTY, F_R_FULL0, F_R_FULL1, F_NBAR };
constexpr uint32_t kFwdAux = 5 * 1024;     // pmax [2][2][128] | rowsum [2][2][128] | rowmax [2][128]

__host__ __device__ inline uint32_t fwd_smem_bytes(int Lk16, const Geo& g) {
  return (uint32_t)(2 * g.nbox * Lk16 * g.rowb + g.nbox * kTileM * g.rowb) + kFwdAux + 256 /*barriers*/ + 1024 /*alignment*/;
}

template <int N, bool kMask>
__device__ __forceinline__ float fwd_max_chunk(uint32_t ts, int c0, int Lk, float mx) {
  uint32_t v[N];
  tmem_ld<N>(ts + c0, v);
  tmem_wait_ld();
#pragma unroll
  for (int j = 0; j < N; ++j)
    if (!kMask || c0 + j < Lk) mx = fmaxf(mx, __uint_as_float(v[j]));
  return mx;
}
template <int N, bool kMask>
__device__ __forceinline__ float fwd_exp_chunk(uint32_t ts, uint32_t tp, int c0, const FwdParams& p, float mxl, uint64_t idx0) {
  uint32_t v[N];
  tmem_ld<N>(ts + c0, v);
  tmem_wait_ld();
  float e[N];
  float sum = 0.f;
#pragma unroll
  for (int j = 0; j < N; ++j) {
    e[j] = psg_ex2_approx(fmaf(__uint_as_float(v[j]), p.scale_l2, -mxl));
    if (kMask && c0 + j >= p.Lk) e[j] = 0.f;
    sum += e[j];
  }
  if (p.drop.thr) drop_mask<N>(e, p.drop, idx0 + (uint64_t)c0);
  uint32_t pk[N / 2];
#pragma unroll
  for (int j = 0; j < N / 2; ++j) pk[j] = pack_bf16(e[2 * j], e[2 * j + 1]);
  tmem_st<N / 2>(tp + (c0 >> 1), pk);
  return sum;
}
// N fp32 accumulator columns of this thread's row -> * f -> bf16 -> global.  A lane's 16 columns are one 32-byte sector: with
// 32-byte-aligned rows (`wide`) it goes out as ONE 256-bit store (16-byte stores made every sector two LSU wavefronts, and the
// drains were bound by exactly that).
__device__ __forceinline__ void st_global_256(void* ptr, const uint32_t (&v)[8]) {
  asm volatile("st.global.v8.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};" ::"l"(ptr), "r"(v[0]), "r"(v[1]), "r"(v[2]), "r"(v[3]), "r"(v[4]),
               "r"(v[5]), "r"(v[6]), "r"(v[7])
               : "memory");
}
template <int N>
__device__ __forceinline__ void store_chunk(uint32_t taddr, int c0, float f, __nv_bfloat16* row_ptr, bool valid, bool wide) {
  uint32_t v[N];
  tmem_ld<N>(taddr + c0, v);
  tmem_wait_ld();
  if (valid) {
#pragma unroll
    for (int g16 = 0; g16 < N / 16; ++g16) {
      uint32_t o[8];
#pragma unroll
      for (int j = 0; j < 8; ++j) o[j] = pack_bf16(__uint_as_float(v[16 * g16 + 2 * j]) * f, __uint_as_float(v[16 * g16 + 2 * j + 1]) * f);
      __nv_bfloat16* dst = row_ptr + c0 + 16 * g16;
      if (wide) {
        st_global_256(dst, o);
      } else {
        *reinterpret_cast<uint4*>(dst) = make_uint4(o[0], o[1], o[2], o[3]);
        *reinterpret_cast<uint4*>(dst + 8) = make_uint4(o[4], o[5], o[6], o[7]);
      }
    }
  }
}
// columns [0, n) of a drained accumulator, 32-column chunks first, first + 2 * step, ... (+ a 16-column tail), for one of `nw` warps
__device__ __forceinline__ void store_cols(uint32_t taddr, int n, int first, int nw, float f, __nv_bfloat16* row_ptr, bool valid, bool wide) {
  const int nch = (n + 31) >> 5;
#pragma unroll 1
  for (int c = first; c < nch; c += nw) {
    if (32 * c + 32 <= n) store_chunk<32>(taddr, 32 * c, f, row_ptr, valid, wide);
    else store_chunk<16>(taddr, 32 * c, f, row_ptr, valid, wide);
  }
}
__device__ __forceinline__ void pair_bar_sync(int q) { asm volatile("bar.sync %0, 64;" ::"r"(1 + q) : "memory"); }   // the two warps of a quarter

__global__ void __launch_bounds__(kThreads, 1) fwd_kernel(const __grid_constant__ FwdParams p) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t smem = (smem_u32(smem_raw) + 1023u) & ~1023u;
  uint8_t* smem_gen = smem_raw + (smem - smem_u32(smem_raw));
  const Geo g = p.g;
  const uint32_t kvbox = (uint32_t)p.Lk16 * g.rowb, qbox = (uint32_t)kTileM * g.rowb;
  const uint32_t sK = smem, sV = sK + g.nbox * kvbox, sQ = sV + g.nbox * kvbox;
  const uint32_t aux = sQ + g.nbox * qbox;
  float* pmax = reinterpret_cast<float*>(smem_gen + (aux - smem));      // [tile parity][half][128]
  float* rowsum = pmax + 4 * kTileM;                                    // [tile parity][half][128]
  float* rowmax = rowsum + 4 * kTileM;                                  // [tile parity][128]
  const uint32_t bars = aux + kFwdAux;
  const uint32_t tmem_slot = bars + 8 * F_NBAR;
#define FBAR(i) (bars + 8u * (i))
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

  if (threadIdx.x == 0) {
    prefetch_tmap(&p.tm_q);
    prefetch_tmap(&p.tm_k);
    prefetch_tmap(&p.tm_v);
    for (int i = 0; i < F_NBAR; ++i)
      mbar_init(FBAR(i), (i == F_P_FULL || i == F_R_FULL0 || i == F_R_FULL1) ? 8 : (i == F_O_EMPTY ? 4 : 1));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 1) tmem_alloc(tmem_slot, p.tmem_cols);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *reinterpret_cast<volatile uint32_t*>(smem_gen + (tmem_slot - smem));
  const int ksteps_qk = p.hd / 16, ksteps_pv = p.Lk16 / 16;

  if (warp == 0) {
    if (lane == 0) {
      // ===== TMA producer: K(u), Q(u,0), V(u), Q(u,1), ... =====
      uint32_t it = 0, un = 0;
      for (Walk w(p.n_mt, p.units); w.valid(); w.advance(), ++it) {
        const int b = w.u / p.H, h = w.u - b * p.H;
        if (w.first()) {
          mbar_wait(FBAR(F_K_EMPTY), (un & 1) ^ 1);
          mbar_expect_tx(FBAR(F_K_FULL), g.nbox * kvbox);
          for (int j = 0; j < g.nbox; ++j) tma_load_4d(sK + j * kvbox, &p.tm_k, FBAR(F_K_FULL), j * g.W, h, 0, b);
        }
        mbar_wait(FBAR(F_Q_EMPTY), (it & 1) ^ 1);
        mbar_expect_tx(FBAR(F_Q_FULL), g.nbox * qbox);
        for (int j = 0; j < g.nbox; ++j) tma_load_4d(sQ + j * qbox, &p.tm_q, FBAR(F_Q_FULL), j * g.W, h, w.t * kTileM, b);
        if (w.first()) {
          mbar_wait(FBAR(F_V_EMPTY), (un & 1) ^ 1);
          mbar_expect_tx(FBAR(F_V_FULL), g.nbox * kvbox);
          for (int j = 0; j < g.nbox; ++j) tma_load_4d(sV + j * kvbox, &p.tm_v, FBAR(F_V_FULL), j * g.W, h, 0, b);
        }
        if (w.last()) ++un;
      }
    }
    __syncwarp();
  } else if (warp == 1) {
    if (lane == 0) {
      // ===== MMA issuer: S(0); then per tile: [softmax(i-1) done] S(i), PV(i-1) =====
      const uint32_t idesc_s = make_idesc(p.Lk16, 0, 0), idesc_o = make_idesc(p.hd, 0, 1);
      uint32_t it = 0, un = 0, prev_un = 0;
      bool prev_first = false, prev_last = false;
      PROF_DECL;
      auto issue_pv = [&](uint32_t j, bool first, bool last, uint32_t unit_ctr) {
        if (first) mbar_wait(FBAR(F_V_FULL), unit_ctr & 1);
        mbar_wait(FBAR(F_O_EMPTY), (j & 1) ^ 1);
        tc_fence_after();
        mma_ts_loop(g, tmem_base + p.col_o, tmem_base + p.col_p, 8, sV, kvbox, 0, ksteps_pv, idesc_o, false);
        umma_commit(FBAR(F_P_EMPTY));
        if (last) umma_commit(FBAR(F_V_EMPTY));
        umma_commit(FBAR(F_O_FULL));
      };
      for (Walk w(p.n_mt, p.units); w.valid(); w.advance(), ++it) {
        if (w.first()) mbar_wait(FBAR(F_K_FULL), un & 1);
        mbar_wait(FBAR(F_Q_FULL), it & 1);
        PROF_LAP(0);
        if (it > 0) mbar_wait(FBAR(F_P_FULL), (it - 1) & 1);      // softmax(i-1) has read S and written P
        PROF_LAP(1);
        tc_fence_after();
        mma_ss_loop(g, tmem_base, sQ, qbox, sK, kvbox, 0, ksteps_qk, idesc_s);
        umma_commit(FBAR(F_Q_EMPTY));
        if (w.last()) umma_commit(FBAR(F_K_EMPTY));
        umma_commit(FBAR(F_S_FULL));
        PROF_LAP(2);
        if (it > 0) issue_pv(it - 1, prev_first, prev_last, prev_un);
        PROF_LAP(3);
        prev_first = w.first(); prev_last = w.last(); prev_un = un;
        if (w.last()) ++un;
      }
      if (it > 0) {
        mbar_wait(FBAR(F_P_FULL), (it - 1) & 1);
        issue_pv(it - 1, prev_first, prev_last, prev_un);
      }
      PROF_FLUSH(0, true);
    }
    __syncwarp();
  } else if (warp < 10) {
    // ===== softmax warps: thread = query row of the tile; the warps of a quarter take the 32-column chunks half, half + 2, ...
    const int q = warp & 3, half = (warp - 2) >> 2, row = q * 32 + lane;
    const uint32_t ts = tmem_base + ((uint32_t)(q * 32) << 16), tp = ts + p.col_p;
    const int nch = (p.Lk16 + 31) >> 5;
    uint32_t it = 0;
    PROF_DECL;
    for (Walk w(p.n_mt, p.units); w.valid(); w.advance(), ++it) {
      const int par = it & 1;
      PROF_LAP(5);
      mbar_wait(FBAR(F_S_FULL), par);
      PROF_LAP(0);
      tc_fence_after();
      // a quarter whose 32 rows all lie past the end of the sequence does nothing (its P / statistics are never used: every
      // accumulator row depends on its own A row only)
      const bool live = w.t * kTileM + q * 32 < p.Lq;
      float mx = -INFINITY;
#pragma unroll 1
      for (int c = half; live && c < nch; c += 2) {
        const int c0 = 32 * c;
        if (c0 + 32 <= p.Lk) mx = fwd_max_chunk<32, false>(ts, c0, p.Lk, mx);
        else if (c0 + 32 <= p.Lk16) mx = fwd_max_chunk<32, true>(ts, c0, p.Lk, mx);
        else mx = fwd_max_chunk<16, true>(ts, c0, p.Lk, mx);
      }
      pmax[(par * 2 + half) * kTileM + row] = mx;
      PROF_LAP(1);
      pair_bar_sync(q);
      mx = fmaxf(mx, pmax[(par * 2 + (half ^ 1)) * kTileM + row]);
      PROF_LAP(2);
      mbar_wait(FBAR(F_P_EMPTY), par ^ 1);                         // PV(i-1) has consumed the previous P
      PROF_LAP(3);
      tc_fence_after();
      const int row_q = w.t * kTileM + row;
      const uint64_t idx0 = ((uint64_t)w.u * p.Lq + (uint64_t)row_q) * (uint64_t)p.Lk;
      const float mxl = mx * p.scale_l2;
      float sum = 0.f;
#pragma unroll 1
      for (int c = half; live && c < nch; c += 2) {
        const int c0 = 32 * c;
        if (c0 + 32 <= p.Lk) sum += fwd_exp_chunk<32, false>(ts, tp, c0, p, mxl, idx0);
        else if (c0 + 32 <= p.Lk16) sum += fwd_exp_chunk<32, true>(ts, tp, c0, p, mxl, idx0);
        else sum += fwd_exp_chunk<16, true>(ts, tp, c0, p, mxl, idx0);
      }
      tmem_wait_st();
      PROF_LAP(4);
      rowsum[(par * 2 + half) * kTileM + row] = sum;
      if (half == 0) rowmax[par * kTileM + row] = mx;
      tc_fence_before();
      __syncwarp();
      if (lane == 0) { mbar_arrive(FBAR(F_P_FULL)); mbar_arrive(FBAR(F_R_FULL0 + par)); }
    }
    PROF_FLUSH(1, warp == 2 && lane == 0);
  } else {
    // ===== epilogue warps: O * (ks / rowsum) -> bf16 -> global; LSE =====
    const int q = warp & 3, row = q * 32 + lane;
    const uint32_t to = tmem_base + ((uint32_t)(q * 32) << 16) + p.col_o;
    uint32_t it = 0;
    PROF_DECL;
    for (Walk w(p.n_mt, p.units); w.valid(); w.advance(), ++it) {
      const int b = w.u / p.H, h = w.u - b * p.H, par = it & 1;
      PROF_LAP(1);
      mbar_wait_backoff(FBAR(F_O_FULL), par);
      PROF_LAP(0);
      mbar_wait(FBAR(F_R_FULL0 + par), (it >> 1) & 1);              // (completed long ago) acquire for the row statistics
      tc_fence_after();
      const float sum = rowsum[(par * 2) * kTileM + row] + rowsum[(par * 2 + 1) * kTileM + row];
      const float f = p.ks / sum;
      const int row_q = w.t * kTileM + row;
      const bool valid = row_q < p.Lq;
      if (p.lse && valid) p.lse[(long long)w.u * p.Lq + row_q] = fmaf(rowmax[par * kTileM + row], p.scale, __logf(sum));
      __nv_bfloat16* orow = p.out + ((long long)b * p.Lq + row_q) * p.ldo + (long long)h * p.hd;
      if (w.t * kTileM + q * 32 < p.Lq) store_cols(to, p.hd, 0, 1, f, orow, valid, p.wide_out != 0);
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(FBAR(F_O_EMPTY));
    }
    PROF_FLUSH(2, warp == 10 && lane == 0);
  }
#undef FBAR
  tc_fence_before();
  __syncthreads();
  if (warp == 1) tmem_dealloc(tmem_base, p.tmem_cols);
}

// =====================================================================================================================
// backward
//   warps 2..9: the elementwise work (two warps per TMEM lane quarter, each a contiguous range of 16-column pieces, two pieces
//   per TMEM round trip); warps 10..13: row / column statistics (delta, LSE) staged one tile / unit AHEAD of their use.
// =====================================================================================================================
struct BwdParams {
  CUtensorMap tm_q, tm_k, tm_v, tm_do;
  const __nv_bfloat16 *o, *dout;
  long long ldo, lddo;
  __nv_bfloat16 *dq, *dk, *dv;
  long long lddq, lddk, lddv;
  const float* lse;
  float* delta;             // [B, H, Lq]: written by delta_kernel, read by the dQ and dK/dV kernels
  int B, H, Lq, Lk, hd, Lq16, Lk16;
  Geo g;
  int n_t, units;           // tiles per unit (query tiles / key tiles)
  float scale, scale_l2, ks;
  Drop drop;
  int col_a, col_b;         // dQ kernel: dP, dQ columns; dK/dV kernel: dP^T, (unused)
  int col_dv, col_dk;
  uint32_t tmem_cols;
  int wide_out;             // dq / dk / dv rows are 32-byte aligned: 256-bit stores
};

// delta[b, h, i] = dO_i . O_i over the head's columns: two lanes per (token, head), half a head each, every load of a lane in
// flight at once; a warp's 16 (token, head) pairs are consecutive heads of consecutive tokens, i.e. contiguous memory.
__global__ void __launch_bounds__(256) delta_kernel(const __nv_bfloat16* __restrict__ o, long long ldo, const __nv_bfloat16* __restrict__ dout,
                                                    long long lddo, float* __restrict__ delta, int B, int H, int Lq, int hd) {
  const long long pair = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 1;      // (token, head)
  const long long npairs = (long long)B * Lq * H;
  const int side = threadIdx.x & 1, half = hd >> 1;
  float s = 0.f;
  long long tok = 0;
  int h = 0;
  if (pair < npairs) {
    tok = pair / H;
    h = (int)(pair - tok * H);
    const __nv_bfloat16* orow = o + tok * ldo + (long long)h * hd + side * half;
    const __nv_bfloat16* drow = dout + tok * lddo + (long long)h * hd + side * half;
#pragma unroll 10
    for (int c = 0; c < half; c += 8) {
      Vec8<__nv_bfloat16> a, d;
      a.load(orow + c);
      d.load(drow + c);
#pragma unroll
      for (int x = 0; x < 8; ++x) s = fmaf(a.v[x], d.v[x], s);
    }
  }
  s += __shfl_xor_sync(0xffffffffu, s, 1);
  if (pair < npairs && side == 0) {
    const long long b = tok / Lq, qi = tok - b * Lq;
    delta[(b * H + h) * Lq + qi] = s;
  }
}

// this warp's contiguous share of `np` 16-column pieces: half 0 gets the first ceil(np / 2)
__device__ __forceinline__ void piece_range(int np, int half, int& lo, int& hi) {
  const int mid = (np + 1) >> 1;
  lo = half ? mid : 0;
  hi = half ? np : mid;
}

// ---------------------------------------------------------------------------------------------------------------------
// dQ: unit = (batch, head): K, V resident; tile = 128 queries (Q, dO tiles streamed).
//   TMEM: S [0, Lk16) | dP [Lk16, 2 Lk16) | dS bf16: piece c (16 keys) packed into the first 8 columns of S's piece c
//         dQ [Lk16, Lk16 + hd) (over dP, dead by then)
// ---------------------------------------------------------------------------------------------------------------------
enum { Q_KV_FULL, Q_KV_EMPTY, Q_QD_FULL, Q_QD_EMPTY, Q_SD_FULL, Q_DS_FULL, Q_DQ_FULL, Q_T_EMPTY, Q_ST_FULL0, Q_ST_FULL1, Q_ST_EMPTY0,
       Q_ST_EMPTY1, Q_NBAR };

__host__ __device__ inline uint32_t dq_smem_bytes(int Lk16, const Geo& g) {
  return (uint32_t)(2 * g.nbox * Lk16 * g.rowb + 2 * g.nbox * kTileM * g.rowb) + 2048 /*lse, delta [2][128]*/ + 256 + 1024;
}

template <int NP>
__device__ __forceinline__ void dq_pieces(uint32_t ts, uint32_t tdp, int c0, const BwdParams& p, float lse2, float delta, uint64_t idx0) {
  uint32_t s[NP][16], d[NP][16];
#pragma unroll
  for (int i = 0; i < NP; ++i) {
    tmem_ld16(ts + c0 + 16 * i, s[i]);
    tmem_ld16(tdp + c0 + 16 * i, d[i]);
  }
  tmem_wait_ld();
#pragma unroll
  for (int i = 0; i < NP; ++i) {
    const int cc = c0 + 16 * i, gk = cc;
    float dpd[16];
#pragma unroll
    for (int j = 0; j < 16; ++j) dpd[j] = __uint_as_float(d[i][j]) * p.ks;
    if (p.drop.thr) drop_mask<16>(dpd, p.drop, idx0 + (uint64_t)gk);
    const bool masked = gk + 16 > p.Lk;
    uint32_t pk[8];
#pragma unroll
    for (int j = 0; j < 16; j += 2) {
      float p0 = psg_ex2_approx(fmaf(__uint_as_float(s[i][j]), p.scale_l2, -lse2));
      float p1 = psg_ex2_approx(fmaf(__uint_as_float(s[i][j + 1]), p.scale_l2, -lse2));
      if (masked && gk + j >= p.Lk) p0 = 0.f;
      if (masked && gk + j + 1 >= p.Lk) p1 = 0.f;
      pk[j >> 1] = pack_bf16(p0 * (dpd[j] - delta), p1 * (dpd[j + 1] - delta));
    }
    tmem_st8(ts + cc, pk);
  }
}

__global__ void __launch_bounds__(kThreads, 1) bwd_dq_kernel(const __grid_constant__ BwdParams p) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t smem = (smem_u32(smem_raw) + 1023u) & ~1023u;
  uint8_t* smem_gen = smem_raw + (smem - smem_u32(smem_raw));
  const Geo g = p.g;
  const uint32_t kvbox = (uint32_t)p.Lk16 * g.rowb, qbox = (uint32_t)kTileM * g.rowb;
  const uint32_t sK = smem, sV = sK + g.nbox * kvbox, sQ = sV + g.nbox * kvbox, sdO = sQ + g.nbox * qbox;
  const uint32_t aux = sdO + g.nbox * qbox;
  float* lse_s = reinterpret_cast<float*>(smem_gen + (aux - smem));      // [tile parity][128] (log2 units)
  float* del_s = lse_s + 2 * kTileM;                                     // [tile parity][128]
  const uint32_t bars = aux + 2048;
  const uint32_t tmem_slot = bars + 8 * Q_NBAR;
#define QBAR(i) (bars + 8u * (i))
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

  if (threadIdx.x == 0) {
    prefetch_tmap(&p.tm_q);
    prefetch_tmap(&p.tm_k);
    prefetch_tmap(&p.tm_v);
    prefetch_tmap(&p.tm_do);
    for (int i = 0; i < Q_NBAR; ++i)
      mbar_init(QBAR(i), (i == Q_DS_FULL || i == Q_T_EMPTY || i == Q_ST_EMPTY0 || i == Q_ST_EMPTY1) ? 8 : ((i == Q_ST_FULL0 || i == Q_ST_FULL1) ? 4 : 1));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 1) tmem_alloc(tmem_slot, p.tmem_cols);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *reinterpret_cast<volatile uint32_t*>(smem_gen + (tmem_slot - smem));
  const int ksteps_hd = p.hd / 16, ksteps_lk = p.Lk16 / 16;

  if (warp == 0) {
    if (lane == 0) {
      uint32_t it = 0, un = 0;
      for (Walk w(p.n_t, p.units); w.valid(); w.advance(), ++it) {
        const int b = w.u / p.H, h = w.u - b * p.H;
        if (w.first()) {
          mbar_wait(QBAR(Q_KV_EMPTY), (un & 1) ^ 1);
          mbar_expect_tx(QBAR(Q_KV_FULL), 2 * g.nbox * kvbox);
          for (int j = 0; j < g.nbox; ++j) tma_load_4d(sK + j * kvbox, &p.tm_k, QBAR(Q_KV_FULL), j * g.W, h, 0, b);
          for (int j = 0; j < g.nbox; ++j) tma_load_4d(sV + j * kvbox, &p.tm_v, QBAR(Q_KV_FULL), j * g.W, h, 0, b);
        }
        mbar_wait(QBAR(Q_QD_EMPTY), (it & 1) ^ 1);
        mbar_expect_tx(QBAR(Q_QD_FULL), 2 * g.nbox * qbox);
        for (int j = 0; j < g.nbox; ++j) tma_load_4d(sQ + j * qbox, &p.tm_q, QBAR(Q_QD_FULL), j * g.W, h, w.t * kTileM, b);
        for (int j = 0; j < g.nbox; ++j) tma_load_4d(sdO + j * qbox, &p.tm_do, QBAR(Q_QD_FULL), j * g.W, h, w.t * kTileM, b);
        if (w.last()) ++un;
      }
    }
    __syncwarp();
  } else if (warp == 1) {
    if (lane == 0) {
      const uint32_t idesc_s = make_idesc(p.Lk16, 0, 0), idesc_q = make_idesc(p.hd, 0, 1);
      uint32_t it = 0, un = 0;
      PROF_DECL;
      for (Walk w(p.n_t, p.units); w.valid(); w.advance(), ++it) {
        if (w.first()) mbar_wait(QBAR(Q_KV_FULL), un & 1);
        mbar_wait(QBAR(Q_QD_FULL), it & 1);
        PROF_LAP(0);
        tc_fence_after();
        mma_ss_loop(g, tmem_base, sQ, qbox, sK, kvbox, 0, ksteps_hd, idesc_s);       // (queued behind the previous tile's dQ MMAs, which read dS out of these columns)
        PROF_LAP(2);
        mbar_wait(QBAR(Q_T_EMPTY), (it & 1) ^ 1);                  // the previous tile's dQ (in dP's columns) has been drained
        PROF_LAP(1);
        tc_fence_after();
        mma_ss_loop(g, tmem_base + p.col_a, sdO, qbox, sV, kvbox, 0, ksteps_hd, idesc_s);
        umma_commit(QBAR(Q_QD_EMPTY));
        umma_commit(QBAR(Q_SD_FULL));
        PROF_LAP(2);
        mbar_wait(QBAR(Q_DS_FULL), it & 1);
        PROF_LAP(3);
        tc_fence_after();
        mma_ts_loop(g, tmem_base + p.col_b, tmem_base, 16, sK, kvbox, 0, ksteps_lk, idesc_q, false);
        if (w.last()) { umma_commit(QBAR(Q_KV_EMPTY)); ++un; }
        umma_commit(QBAR(Q_DQ_FULL));
        PROF_LAP(4);
      }
      PROF_FLUSH(0, true);
    }
    __syncwarp();
  } else if (warp < 10) {
    // ===== compute warps: dS, dQ drain =====
    const int q = warp & 3, half = (warp - 2) >> 2, row = q * 32 + lane;
    const uint32_t ts = tmem_base + ((uint32_t)(q * 32) << 16), tdp = ts + p.col_a, tdq = ts + p.col_b;
    int plo, phi;
    piece_range(ksteps_lk, half, plo, phi);
    uint32_t it = 0;
    PROF_DECL;
    for (Walk w(p.n_t, p.units); w.valid(); w.advance(), ++it) {
      const int b = w.u / p.H, h = w.u - b * p.H, par = it & 1;
      PROF_LAP(4);
      mbar_wait(QBAR(Q_ST_FULL0 + par), (it >> 1) & 1);
      PROF_LAP(0);
      const float lse2 = lse_s[par * kTileM + row], delta = del_s[par * kTileM + row];
      const int row_q = w.t * kTileM + row;
      const uint64_t idx0 = ((uint64_t)w.u * p.Lq + (uint64_t)row_q) * (uint64_t)p.Lk;
      mbar_wait(QBAR(Q_SD_FULL), par);
      PROF_LAP(1);
      tc_fence_after();
      const bool live = w.t * kTileM + q * 32 < p.Lq;      // (a quarter of rows past the sequence end does nothing, see the forward)
      int c = live ? plo : phi;
#pragma unroll 1
      for (; c + 2 <= phi; c += 2) dq_pieces<2>(ts, tdp, 16 * c, p, lse2, delta, idx0);
      if (c < phi) dq_pieces<1>(ts, tdp, 16 * c, p, lse2, delta, idx0);
      tmem_wait_st();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) { mbar_arrive(QBAR(Q_DS_FULL)); mbar_arrive(QBAR(Q_ST_EMPTY0 + par)); }
      PROF_LAP(2);
      mbar_wait(QBAR(Q_DQ_FULL), par);
      PROF_LAP(3);
      tc_fence_after();
      const bool valid = row_q < p.Lq;
      __nv_bfloat16* qrow = p.dq + ((long long)b * p.Lq + row_q) * p.lddq + (long long)h * p.hd;
      if (live) store_cols(tdq, p.hd, half, 2, p.scale, qrow, valid, p.wide_out != 0);
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(QBAR(Q_T_EMPTY));
    }
    PROF_FLUSH(1, warp == 2 && lane == 0);
  } else {
    // ===== statistics warps: LSE (log2 units) and delta (delta_kernel's output) of the NEXT tile's rows
    const int sw = warp - 10;
    uint32_t it = 0;
    PROF_DECL;
    for (Walk w(p.n_t, p.units); w.valid(); w.advance(), ++it) {
      const int par = it & 1;
      PROF_LAP(1);
      mbar_wait_backoff(QBAR(Q_ST_EMPTY0 + par), ((it >> 1) & 1) ^ 1);
      PROF_LAP(0);
      {
        const int r = sw * 32 + lane, qi = w.t * kTileM + r;
        const bool ok = qi < p.Lq;
        const long long rg = (long long)w.u * p.Lq + qi;
        lse_s[par * kTileM + r] = ok ? p.lse[rg] * 1.4426950408889634f : 0.f;
        del_s[par * kTileM + r] = ok ? p.delta[rg] : 0.f;
      }
      __syncwarp();
      if (lane == 0) mbar_arrive(QBAR(Q_ST_FULL0 + par));
    }
    PROF_FLUSH(2, warp == 10 && lane == 0);
  }
#undef QBAR
  tc_fence_before();
  __syncthreads();
  if (warp == 1) tmem_dealloc(tmem_base, p.tmem_cols);
}

// ---------------------------------------------------------------------------------------------------------------------
// dK / dV: unit = (batch, head): Q, dO resident (all Lq16 rows); tile = 128 keys (K, V tiles streamed); per tile the queries
// are walked in blocks of <= 96:
//   TMEM: S^T [0, 96) | dP^T [96, 192) | dV [192, 192 + hd) | dK [192 + hd, 192 + 2 hd)
//         Pd^T / dS^T bf16: piece c (16 queries) packed into the first 8 columns of S^T's / dP^T's piece c
// ---------------------------------------------------------------------------------------------------------------------
enum { D_QD_FULL, D_QD_EMPTY, D_KV_FULL, D_KV_EMPTY, D_ST_FULL, D_PD_FULL, D_ACC_FULL, D_ACC_EMPTY, D_STAT_FULL0, D_STAT_FULL1, D_STAT_EMPTY0,
       D_STAT_EMPTY1, D_NBAR };

__host__ __device__ inline uint32_t dkv_smem_bytes(int Lq16, const Geo& g) {
  return (uint32_t)(2 * g.nbox * Lq16 * g.rowb + 2 * g.nbox * kTileM * g.rowb) + 4096 /*lse, delta [2][256] each*/ + 256 + 1024;
}

template <int NP>
__device__ __forceinline__ void dkv_pieces(uint32_t tst, uint32_t tdp, int c0, int q0, const BwdParams& p, const float* lse_s,
                                           const float* del_s, uint64_t unit_row0, int key) {
  uint32_t s[NP][16], d[NP][16];
#pragma unroll
  for (int i = 0; i < NP; ++i) {
    tmem_ld16(tst + c0 + 16 * i, s[i]);
    tmem_ld16(tdp + c0 + 16 * i, d[i]);
  }
  tmem_wait_ld();
  const uint32_t thr16 = p.drop.thr >> 16;
  const bool key_ok = key < p.Lk;
  const bool fast = p.drop.idx32 && ((p.Lk & 1) == 0);            // every (q, *) row starts on an even flat index
#pragma unroll
  for (int i = 0; i < NP; ++i) {
    const int qq = q0 + c0 + 16 * i;          // first query of the piece
    float l2[16], de[16];
#pragma unroll
    for (int j = 0; j < 16; j += 4) {
      const float4 a = *reinterpret_cast<const float4*>(lse_s + qq + j);
      const float4 e = *reinterpret_cast<const float4*>(del_s + qq + j);
      l2[j] = a.x; l2[j + 1] = a.y; l2[j + 2] = a.z; l2[j + 3] = a.w;
      de[j] = e.x; de[j + 1] = e.y; de[j + 2] = e.z; de[j + 3] = e.w;
    }
    float m[16];                              // dropout multiplier: ks or 0
    if (!p.drop.thr) {
#pragma unroll
      for (int j = 0; j < 16; ++j) m[j] = 1.f;
    } else if (fast) {
      // flat index of (q, key) = (unit_row0 + q) * Lk + key: pair index = that >> 1, field = key & 1
      const uint32_t pair0 = (uint32_t)((((unit_row0 + (uint64_t)qq) * (uint64_t)p.Lk) + (uint64_t)key) >> 1);
      const uint32_t step = (uint32_t)p.Lk >> 1;
      const uint32_t sh = (key & 1) ? 16u : 0u;
#pragma unroll
      for (int j = 0; j < 16; ++j) {
        const uint32_t hsh = mix32((pair0 + (uint32_t)j * step) ^ p.drop.seedmix);
        m[j] = (((hsh >> sh) & 0xFFFFu) >= thr16) ? p.ks : 0.f;
      }
    } else {
#pragma unroll
      for (int j = 0; j < 16; ++j)
        m[j] = psg_drop_keep(p.drop.seed, (unit_row0 + (uint64_t)(qq + j)) * (uint64_t)p.Lk + (uint64_t)key, p.drop.thr) ? p.ks : 0.f;
    }
    uint32_t pp[8], pd[8];
#pragma unroll
    for (int j = 0; j < 16; j += 2) {
      float pr0 = psg_ex2_approx(fmaf(__uint_as_float(s[i][j]), p.scale_l2, -l2[j]));          // lse_s = +inf past Lq: 0
      float pr1 = psg_ex2_approx(fmaf(__uint_as_float(s[i][j + 1]), p.scale_l2, -l2[j + 1]));
      if (!key_ok) { pr0 = 0.f; pr1 = 0.f; }
      pp[j >> 1] = pack_bf16(pr0 * m[j], pr1 * m[j + 1]);
      pd[j >> 1] = pack_bf16(pr0 * fmaf(__uint_as_float(d[i][j]), m[j], -de[j]), pr1 * fmaf(__uint_as_float(d[i][j + 1]), m[j + 1], -de[j + 1]));
    }
    tmem_st8(tst + c0 + 16 * i, pp);
    tmem_st8(tdp + c0 + 16 * i, pd);
  }
}

__global__ void __launch_bounds__(kThreads, 1) bwd_dkv_kernel(const __grid_constant__ BwdParams p) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t smem = (smem_u32(smem_raw) + 1023u) & ~1023u;
  uint8_t* smem_gen = smem_raw + (smem - smem_u32(smem_raw));
  const Geo g = p.g;
  const uint32_t qabox = (uint32_t)p.Lq16 * g.rowb, kbox = (uint32_t)kTileM * g.rowb;
  const uint32_t sQ = smem, sdO = sQ + g.nbox * qabox, sK = sdO + g.nbox * qabox, sV = sK + g.nbox * kbox;
  const uint32_t aux = sV + g.nbox * kbox;
  float* lse_s = reinterpret_cast<float*>(smem_gen + (aux - smem));      // [unit parity][256] (log2 units, +inf past Lq)
  float* del_s = lse_s + 512;                                            // [unit parity][256]
  const uint32_t bars = aux + 4096;
  const uint32_t tmem_slot = bars + 8 * D_NBAR;
#define DBAR(i) (bars + 8u * (i))
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

  if (threadIdx.x == 0) {
    prefetch_tmap(&p.tm_q);
    prefetch_tmap(&p.tm_k);
    prefetch_tmap(&p.tm_v);
    prefetch_tmap(&p.tm_do);
    for (int i = 0; i < D_NBAR; ++i)
      mbar_init(DBAR(i), (i == D_PD_FULL || i == D_ACC_EMPTY || i == D_STAT_EMPTY0 || i == D_STAT_EMPTY1) ? 8
                             : ((i == D_STAT_FULL0 || i == D_STAT_FULL1) ? 4 : 1));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 1) tmem_alloc(tmem_slot, p.tmem_cols);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *reinterpret_cast<volatile uint32_t*>(smem_gen + (tmem_slot - smem));
  const int ksteps_hd = p.hd / 16;
  const int nblk = (p.Lq16 + kQBlk - 1) / kQBlk;

  if (warp == 0) {
    if (lane == 0) {
      uint32_t it = 0, un = 0;
      for (Walk w(p.n_t, p.units); w.valid(); w.advance(), ++it) {
        const int b = w.u / p.H, h = w.u - b * p.H;
        if (w.first()) {
          mbar_wait(DBAR(D_QD_EMPTY), (un & 1) ^ 1);
          mbar_expect_tx(DBAR(D_QD_FULL), 2 * g.nbox * qabox);
          for (int j = 0; j < g.nbox; ++j) tma_load_4d(sQ + j * qabox, &p.tm_q, DBAR(D_QD_FULL), j * g.W, h, 0, b);
          for (int j = 0; j < g.nbox; ++j) tma_load_4d(sdO + j * qabox, &p.tm_do, DBAR(D_QD_FULL), j * g.W, h, 0, b);
        }
        mbar_wait(DBAR(D_KV_EMPTY), (it & 1) ^ 1);
        mbar_expect_tx(DBAR(D_KV_FULL), 2 * g.nbox * kbox);
        for (int j = 0; j < g.nbox; ++j) tma_load_4d(sK + j * kbox, &p.tm_k, DBAR(D_KV_FULL), j * g.W, h, w.t * kTileM, b);
        for (int j = 0; j < g.nbox; ++j) tma_load_4d(sV + j * kbox, &p.tm_v, DBAR(D_KV_FULL), j * g.W, h, w.t * kTileM, b);
        if (w.last()) ++un;
      }
    }
    __syncwarp();
  } else if (warp == 1) {
    if (lane == 0) {
      const uint32_t idesc_acc = make_idesc(p.hd, 0, 1);
      uint32_t it = 0, un = 0, blk = 0;         // blk: global block counter (phases of ST_FULL / PD_FULL)
      PROF_DECL;
      for (Walk w(p.n_t, p.units); w.valid(); w.advance(), ++it) {
        if (w.first()) mbar_wait(DBAR(D_QD_FULL), un & 1);
        mbar_wait(DBAR(D_KV_FULL), it & 1);
        PROF_LAP(0);
        tc_fence_after();
        for (int jb = 0; jb < nblk; ++jb, ++blk) {
          const int q0 = jb * kQBlk, nq = min(kQBlk, p.Lq16 - q0);
          const uint32_t idesc_st = make_idesc(nq, 0, 0);
          // S^T = K_t Q_blk^T, dP^T = V_t dO_blk^T (queued behind the previous block's dV / dK MMAs, which read these columns)
          mma_ss_loop(g, tmem_base, sK, kbox, sQ, qabox, q0, ksteps_hd, idesc_st);
          mma_ss_loop(g, tmem_base + p.col_a, sV, kbox, sdO, qabox, q0, ksteps_hd, idesc_st);
          if (jb == nblk - 1) umma_commit(DBAR(D_KV_EMPTY));        // the K / V tile is free once these have run
          umma_commit(DBAR(D_ST_FULL));
          PROF_LAP(1);
          mbar_wait(DBAR(D_PD_FULL), blk & 1);
          if (jb == 0) mbar_wait(DBAR(D_ACC_EMPTY), (it & 1) ^ 1);  // the previous tile's dV / dK have been drained
          PROF_LAP(2);
          tc_fence_after();
          mma_ts_loop(g, tmem_base + p.col_dv, tmem_base, 16, sdO, qabox, q0, nq / 16, idesc_acc, jb > 0);
          mma_ts_loop(g, tmem_base + p.col_dk, tmem_base + p.col_a, 16, sQ, qabox, q0, nq / 16, idesc_acc, jb > 0);
          PROF_LAP(3);
        }
        if (w.last()) { umma_commit(DBAR(D_QD_EMPTY)); ++un; }
        umma_commit(DBAR(D_ACC_FULL));
      }
      PROF_FLUSH(0, true);
    }
    __syncwarp();
  } else if (warp < 10) {
    const int q = warp & 3, half = (warp - 2) >> 2, row = q * 32 + lane;
    const uint32_t tst = tmem_base + ((uint32_t)(q * 32) << 16), tdp = tst + p.col_a;
    uint32_t it = 0, blk = 0, un = 0;
    PROF_DECL;
    for (Walk w(p.n_t, p.units); w.valid(); w.advance(), ++it) {
      const int b = w.u / p.H, h = w.u - b * p.H, upar = un & 1;
      PROF_LAP(4);
      if (w.first()) mbar_wait(DBAR(D_STAT_FULL0 + upar), (un >> 1) & 1);
      PROF_LAP(0);
      const float* lse_u = lse_s + upar * 256;
      const float* del_u = del_s + upar * 256;
      const int key = w.t * kTileM + row;
      const bool live = w.t * kTileM + q * 32 < p.Lk;      // (a quarter of keys past the sequence end does nothing, see the forward)
      const uint64_t unit_row0 = (uint64_t)w.u * (uint64_t)p.Lq;
      for (int jb = 0; jb < nblk; ++jb, ++blk) {
        const int q0 = jb * kQBlk, nq = min(kQBlk, p.Lq16 - q0);
        int plo, phi;
        piece_range(nq / 16, half, plo, phi);
        mbar_wait(DBAR(D_ST_FULL), blk & 1);
        PROF_LAP(1);
        tc_fence_after();
        int c = live ? plo : phi;
#pragma unroll 1
        for (; c + 2 <= phi; c += 2) dkv_pieces<2>(tst, tdp, 16 * c, q0, p, lse_u, del_u, unit_row0, key);
        if (c < phi) dkv_pieces<1>(tst, tdp, 16 * c, q0, p, lse_u, del_u, unit_row0, key);
        tmem_wait_st();
        tc_fence_before();
        __syncwarp();
        if (lane == 0) {
          mbar_arrive(DBAR(D_PD_FULL));
          if (w.last() && jb == nblk - 1) mbar_arrive(DBAR(D_STAT_EMPTY0 + upar));
        }
        PROF_LAP(2);
      }
      mbar_wait(DBAR(D_ACC_FULL), it & 1);
      PROF_LAP(3);
      tc_fence_after();
      const bool valid = key < p.Lk;
      __nv_bfloat16* vrow = p.dv + ((long long)b * p.Lk + key) * p.lddv + (long long)h * p.hd;
      __nv_bfloat16* krow = p.dk + ((long long)b * p.Lk + key) * p.lddk + (long long)h * p.hd;
      if (live) {
        store_cols(tst + p.col_dv, p.hd, half, 2, 1.f, vrow, valid, p.wide_out != 0);
        store_cols(tst + p.col_dk, p.hd, half ^ 1, 2, p.scale, krow, valid, p.wide_out != 0);
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(DBAR(D_ACC_EMPTY));
      if (w.last()) ++un;
    }
    PROF_FLUSH(1, warp == 2 && lane == 0);
  } else {
    // ===== statistics warps: LSE (log2 units, +inf past Lq) and delta of the NEXT unit's queries
    uint32_t un = 0;
    for (Walk w(p.n_t, p.units); w.valid(); w.advance()) {
      if (!w.first()) continue;
      const int upar = un & 1;
      mbar_wait_backoff(DBAR(D_STAT_EMPTY0 + upar), ((un >> 1) & 1) ^ 1);
      for (int i = threadIdx.x - 320; i < 256; i += 128) {
        const bool ok = i < p.Lq;
        lse_s[upar * 256 + i] = ok ? p.lse[(long long)w.u * p.Lq + i] * 1.4426950408889634f : INFINITY;
        del_s[upar * 256 + i] = ok ? p.delta[(long long)w.u * p.Lq + i] : 0.f;
      }
      __syncwarp();
      if (lane == 0) mbar_arrive(DBAR(D_STAT_FULL0 + upar));
      ++un;
    }
  }
#undef DBAR
  tc_fence_before();
  __syncthreads();
  if (warp == 1) tmem_dealloc(tmem_base, p.tmem_cols);
}


// ---------------------------------------------------------------------------------------------------------------------
// host side
// ---------------------------------------------------------------------------------------------------------------------
typedef CUresult (*PFN_encodeTiled)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                    const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                    CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
static PFN_encodeTiled g_encode_tiled = nullptr;

static int load_driver_fns() {
  if (g_encode_tiled) return PSG_OK;
  cudaDriverEntryPointQueryResult qres;
  void* fn = nullptr;
  cudaError_t e = cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qres);
  if (e != cudaSuccess || qres != cudaDriverEntryPointSuccess || !fn) {
    psg_set_error("cuTensorMapEncodeTiled not available: %s", cudaGetErrorString(e));
    return PSG_ERR_CUDA;
  }
  g_encode_tiled = (PFN_encodeTiled)fn;
  return PSG_OK;
}

static Geo make_geo(int hd) {
  Geo g;
  g.W = (hd % 64 == 0) ? 64 : 32;
  g.rowb = g.W * 2;
  g.nbox = (hd + g.W - 1) / g.W;
  g.kshift = (g.W == 64) ? 2 : 1;
  g.swz = (g.W == 64) ? 2u : 4u;
  return g;
}

// token-major [B * L, ld] bf16 with head h at columns [h * hd, (h + 1) * hd): dims (hd, H, L, B), box (W, 1, box_rows, 1)
static int make_map(CUtensorMap* tm, const void* ptr, long long ld, int B, int H, int L, int hd, const Geo& g, int box_rows) {
  cuuint64_t dims[4] = {(cuuint64_t)hd, (cuuint64_t)H, (cuuint64_t)L, (cuuint64_t)B};
  cuuint64_t strides[3] = {(cuuint64_t)hd * 2, (cuuint64_t)ld * 2, (cuuint64_t)ld * 2 * (cuuint64_t)L};
  cuuint32_t box[4] = {(cuuint32_t)g.W, 1, (cuuint32_t)box_rows, 1};
  cuuint32_t estr[4] = {1, 1, 1, 1};
  CUresult r = g_encode_tiled(tm, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, const_cast<void*>(ptr), dims, strides, box, estr,
                              CU_TENSOR_MAP_INTERLEAVE_NONE, g.W == 64 ? CU_TENSOR_MAP_SWIZZLE_128B : CU_TENSOR_MAP_SWIZZLE_64B,
                              CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    psg_set_error("psg_attn_umma: cuTensorMapEncodeTiled failed (%d): B=%d H=%d L=%d hd=%d ld=%lld box_rows=%d ptr=%p", (int)r, B, H, L,
                  hd, ld, box_rows, ptr);
    return PSG_ERR_CUDA;
  }
  return PSG_OK;
}

static Drop make_drop(unsigned long long seed, float drop_p, int B, int H, int Lq, int Lk) {
  Drop d;
  d.seed = seed;
  d.thr = drop_p > 0.f ? (unsigned int)((double)drop_p * 4294967296.0) : 0u;
  d.seedmix = (unsigned int)seed ^ ((unsigned int)(seed >> 32) * 0x85EBCA6Bu);
  d.idx32 = ((double)B * H * Lq * Lk < 8.5e9) ? 1 : 0;       // every pair index (flat index / 2) below 2^32
  return d;
}

static uint32_t pow2_cols(int n) { return n <= 32 ? 32u : n <= 64 ? 64u : n <= 128 ? 128u : n <= 256 ? 256u : 512u; }
static int round16(int x) { return (x + 15) & ~15; }

static int g_enabled = 1;

// the problems the tcgen05 kernels take: sequences that make 128-row tiles worthwhile and operands that fit one SM
static bool problem_ok(int B, int H, int Lq, int Lk, int hd) {
  if (B <= 0 || H <= 0 || hd < 16 || hd % 16 != 0 || hd > 256) return false;
  if (Lq < 65 || Lq > 256 || Lk < 1 || Lk > 256) return false;
  const Geo g = make_geo(hd);
  const int Lq16 = round16(Lq), Lk16 = round16(Lk);
  if (Lk16 + round16(Lk16 / 2) + hd > 512) return false;               // forward: S | P | O
  if (2 * Lk16 > 512 || Lk16 + hd > 512) return false;                 // dQ: S | dP (dQ over dP)
  if (2 * kQBlk + 2 * hd > 512) return false;                          // dK/dV: S^T | dP^T | dV | dK
  if (fwd_smem_bytes(Lk16, g) > kSmemLimit || dq_smem_bytes(Lk16, g) > kSmemLimit || dkv_smem_bytes(Lq16, g) > kSmemLimit) return false;
  return true;
}

template <typename Kern>
static int configure(Kern kern, bool& done, const char* name) {
  if (done) return PSG_OK;
  cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kSmemLimit);
  if (e != cudaSuccess) { psg_set_error("%s: cudaFuncSetAttribute: %s", name, cudaGetErrorString(e)); return PSG_ERR_CUDA; }
  done = true;
  return PSG_OK;
}

}  // namespace uattn

extern "C" {

// Test / measurement hook: 0 = psg_attn_fused_* never route to the tcgen05 kernels, 1 = wherever psg_attn_umma_ok (default).
// Returns the previous value.
int psg_attn_umma_enable(int on) {
  const int prev = uattn::g_enabled;
  if (on == 0 || on == 1) uattn::g_enabled = on;
  return prev;
}

int psg_attn_umma_ok(int B, int H, int Lq, int Lk, int hd) { return uattn::g_enabled && uattn::problem_ok(B, H, Lq, Lk, hd) ? 1 : 0; }

// Test hook: 1 if any mbarrier wait of these kernels timed out since the last call (synchronises the device).
int psg_attn_umma_timeout_flag() {
  int v = 0, zero = 0;
  cudaMemcpyFromSymbol(&v, uattn::g_timeout_flag, sizeof(int));
  cudaMemcpyToSymbol(uattn::g_timeout_flag, &zero, sizeof(int));
  return v;
}

// Tuning aid: 1 = psg_attn_umma_bwd launches only its first kernel (dQ), so that its cycle totals can be read back.
static int g_prof_only = 0;
int psg_attn_umma_prof_only(int on) { g_prof_only = on; return PSG_OK; }

// Tuning aid (library built with -DUATTN_PROF): copies the per-CTA, per-role cycle totals of the last kernel (160 x 32 int64) to
// `out`; returns PSG_ERR_UNSUPPORTED in a normal build.
int psg_attn_umma_prof(long long* out) {
#ifdef UATTN_PROF
  cudaMemcpyFromSymbol(out, uattn::g_prof, sizeof(long long) * 160 * 32);
  return PSG_OK;
#else
  (void)out;
  psg_set_error("psg_attn_umma_prof: library built without UATTN_PROF");
  return PSG_ERR_UNSUPPORTED;
#endif
}

int psg_attn_umma_fwd(const void* q, long long ldq, const void* k, long long ldk, const void* v, long long ldv, void* o, long long ldo,
                      float* lse, int B, int H, int Lq, int Lk, int hd, float scale, unsigned long long drop_seed, float drop_p,
                      void* stream) {
  using namespace uattn;
  PSG_CHECK_ARG(q && k && v && o, "psg_attn_umma_fwd: null pointer");
  PSG_CHECK_ARG(problem_ok(B, H, Lq, Lk, hd), "psg_attn_umma_fwd: unsupported problem B=%d H=%d Lq=%d Lk=%d hd=%d", B, H, Lq, Lk, hd);
  PSG_CHECK_ARG(drop_p >= 0.f && drop_p < 1.f && scale > 0.f, "psg_attn_umma_fwd: bad dropout probability or scale");
  PSG_CHECK_ARG(ldq % 8 == 0 && ldk % 8 == 0 && ldv % 8 == 0 && ldo % 8 == 0 && ((uintptr_t)q % 16 == 0) && ((uintptr_t)k % 16 == 0) &&
                    ((uintptr_t)v % 16 == 0) && ((uintptr_t)o % 16 == 0),
                "psg_attn_umma_fwd: pitches/pointers must be 16B aligned");
  int rc = load_driver_fns();
  if (rc) return rc;
  FwdParams p;
  memset(&p, 0, sizeof(p));
  p.g = make_geo(hd);
  p.B = B; p.H = H; p.Lq = Lq; p.Lk = Lk; p.hd = hd; p.Lk16 = round16(Lk);
  if ((rc = make_map(&p.tm_q, q, ldq, B, H, Lq, hd, p.g, kTileM))) return rc;
  if ((rc = make_map(&p.tm_k, k, ldk, B, H, Lk, hd, p.g, p.Lk16))) return rc;
  if ((rc = make_map(&p.tm_v, v, ldv, B, H, Lk, hd, p.g, p.Lk16))) return rc;
  p.out = (__nv_bfloat16*)o; p.ldo = ldo; p.lse = lse;
  p.n_mt = (Lq + kTileM - 1) / kTileM;
  p.units = B * H;
  p.scale = scale; p.scale_l2 = scale * 1.4426950408889634f;
  p.drop = make_drop(drop_seed, drop_p, B, H, Lq, Lk);
  p.ks = drop_p > 0.f ? 1.f / (1.f - drop_p) : 1.f;
  p.col_p = p.Lk16;
  p.col_o = p.Lk16 + round16(p.Lk16 / 2);
  p.tmem_cols = pow2_cols(p.col_o + hd);
  p.wide_out = (ldo % 16 == 0 && (uintptr_t)o % 32 == 0) ? 1 : 0;
  static bool done = false;
  if ((rc = configure(fwd_kernel, done, "psg_attn_umma_fwd"))) return rc;
  const int grid = p.units < psg_num_sms() ? p.units : psg_num_sms();
  fwd_kernel<<<grid, kThreads, fwd_smem_bytes(p.Lk16, p.g), (cudaStream_t)stream>>>(p);
  PSG_CHECK_LAUNCH("psg_attn_umma_fwd");
  return PSG_OK;
}

int psg_attn_umma_bwd(const void* q, long long ldq, const void* k, long long ldk, const void* v, long long ldv, const void* o,
                      long long ldo, const void* dout, long long lddo, const float* lse, float* delta, void* dq, long long lddq,
                      void* dk, long long lddk, void* dv, long long lddv, int B, int H, int Lq, int Lk, int hd, float scale,
                      unsigned long long drop_seed, float drop_p, void* stream) {
  using namespace uattn;
  PSG_CHECK_ARG(q && k && v && o && dout && lse && delta && dq && dk && dv, "psg_attn_umma_bwd: null pointer");
  PSG_CHECK_ARG(problem_ok(B, H, Lq, Lk, hd), "psg_attn_umma_bwd: unsupported problem B=%d H=%d Lq=%d Lk=%d hd=%d", B, H, Lq, Lk, hd);
  PSG_CHECK_ARG(drop_p >= 0.f && drop_p < 1.f && scale > 0.f, "psg_attn_umma_bwd: bad dropout probability or scale");
  PSG_CHECK_ARG(ldq % 8 == 0 && ldk % 8 == 0 && ldv % 8 == 0 && ldo % 8 == 0 && lddo % 8 == 0 && lddq % 8 == 0 && lddk % 8 == 0 &&
                    lddv % 8 == 0 && ((uintptr_t)q % 16 == 0) && ((uintptr_t)k % 16 == 0) && ((uintptr_t)v % 16 == 0) &&
                    ((uintptr_t)o % 16 == 0) && ((uintptr_t)dout % 16 == 0) && ((uintptr_t)dq % 16 == 0) && ((uintptr_t)dk % 16 == 0) &&
                    ((uintptr_t)dv % 16 == 0),
                "psg_attn_umma_bwd: pitches/pointers must be 16B aligned");
  int rc = load_driver_fns();
  if (rc) return rc;
  BwdParams p;
  memset(&p, 0, sizeof(p));
  p.g = make_geo(hd);
  p.B = B; p.H = H; p.Lq = Lq; p.Lk = Lk; p.hd = hd; p.Lq16 = round16(Lq); p.Lk16 = round16(Lk);
  p.o = (const __nv_bfloat16*)o; p.dout = (const __nv_bfloat16*)dout; p.ldo = ldo; p.lddo = lddo;
  p.dq = (__nv_bfloat16*)dq; p.dk = (__nv_bfloat16*)dk; p.dv = (__nv_bfloat16*)dv; p.lddq = lddq; p.lddk = lddk; p.lddv = lddv;
  p.lse = lse; p.delta = delta;
  p.units = B * H;
  p.scale = scale; p.scale_l2 = scale * 1.4426950408889634f;
  p.drop = make_drop(drop_seed, drop_p, B, H, Lq, Lk);
  p.ks = drop_p > 0.f ? 1.f / (1.f - drop_p) : 1.f;
  p.wide_out = (lddq % 16 == 0 && lddk % 16 == 0 && lddv % 16 == 0 && (uintptr_t)dq % 32 == 0 && (uintptr_t)dk % 32 == 0 && (uintptr_t)dv % 32 == 0) ? 1 : 0;
  const int grid = p.units < psg_num_sms() ? p.units : psg_num_sms();
  cudaStream_t st = (cudaStream_t)stream;
  static bool done1 = false, done2 = false;
  if ((rc = configure(bwd_dq_kernel, done1, "psg_attn_umma_bwd"))) return rc;
  if ((rc = configure(bwd_dkv_kernel, done2, "psg_attn_umma_bwd"))) return rc;
  // ---- delta = rowsum(dO o O), then dQ ----
  {
    const long long threads = 2LL * B * Lq * H;
    delta_kernel<<<(unsigned)((threads + 255) / 256), 256, 0, st>>>(p.o, ldo, p.dout, lddo, delta, B, H, Lq, hd);
    PSG_CHECK_LAUNCH("psg_attn_umma_bwd");
  }
  if ((rc = make_map(&p.tm_q, q, ldq, B, H, Lq, hd, p.g, kTileM))) return rc;
  if ((rc = make_map(&p.tm_do, dout, lddo, B, H, Lq, hd, p.g, kTileM))) return rc;
  if ((rc = make_map(&p.tm_k, k, ldk, B, H, Lk, hd, p.g, p.Lk16))) return rc;
  if ((rc = make_map(&p.tm_v, v, ldv, B, H, Lk, hd, p.g, p.Lk16))) return rc;
  p.n_t = (Lq + kTileM - 1) / kTileM;
  p.col_a = p.Lk16;
  p.col_b = p.Lk16;
  p.tmem_cols = pow2_cols((2 * p.Lk16 > p.Lk16 + hd) ? 2 * p.Lk16 : p.Lk16 + hd);
  bwd_dq_kernel<<<grid, kThreads, dq_smem_bytes(p.Lk16, p.g), st>>>(p);
  PSG_CHECK_LAUNCH("psg_attn_umma_bwd");
  if (g_prof_only) return PSG_OK;
  // ---- dK / dV ----
  if ((rc = make_map(&p.tm_q, q, ldq, B, H, Lq, hd, p.g, p.Lq16))) return rc;
  if ((rc = make_map(&p.tm_do, dout, lddo, B, H, Lq, hd, p.g, p.Lq16))) return rc;
  if ((rc = make_map(&p.tm_k, k, ldk, B, H, Lk, hd, p.g, kTileM))) return rc;
  if ((rc = make_map(&p.tm_v, v, ldv, B, H, Lk, hd, p.g, kTileM))) return rc;
  p.n_t = (Lk + kTileM - 1) / kTileM;
  p.col_a = kQBlk;
  p.col_dv = 2 * kQBlk;
  p.col_dk = 2 * kQBlk + hd;
  p.tmem_cols = pow2_cols(2 * kQBlk + 2 * hd);
  bwd_dkv_kernel<<<grid, kThreads, dkv_smem_bytes(p.Lq16, p.g), st>>>(p);
  PSG_CHECK_LAUNCH("psg_attn_umma_bwd");
  return PSG_OK;
}

}  // extern "C"
