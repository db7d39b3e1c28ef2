// tcgen05 / TMEM / TMA GEMM + implicit-GEMM convolution engine for sm_100a (bf16 in, fp32 accumulate).
//
// One warp-specialised persistent kernel, three roles (320 threads):
//   warp 0      TMA producer  : cp.async.bulk.tensor (tiled 2D, or im2col 4D over NHWC) -> 128B-swizzled smem ring
//   warp 1      MMA issuer    : one thread issues tcgen05.mma (M=128, N=kBlockN, K=16) into a TMEM accumulator
//   warps 2..9  epilogue      : tcgen05.ld TMEM -> registers -> fused epilogue (gemm_epilogue.cuh) -> global
//
// kMode 0 ("TN"):  A K-major (tiled matrix or im2col pixels), B K-major matrix.   fprop / dgrad / linears.
// kMode 1 ("NT"):  A MN-major matrix, B MN-major (tiled matrix or im2col pixels). wgrad (reduction over rows).
// kMode 2 ("TT"):  A as kMode 0, B MN-major: the weight matrix as stored ([N_w, K_w] row-major, conv weights
//                  [Cout][tap][Cin]) read transposed -- dgrad needs no transposed weight copy.
//
// kCl == 2 (CTA pairs, tcgen05 cta_group::2): two CTAs of a thread-block cluster compute ONE 256*kMT-row tile: each holds
// its own 128*kMT rows of A and HALF of the B tile, the leader CTA (rank 0) issues tcgen05.mma.cta_group::2 (M = 256)
// which reads both CTAs' shared memory and writes both CTAs' TMEM.  Per 128x256x64 block a CTA then moves 32+32 KB
// through its shared memory instead of 48+48 KB: measured, the single-CTA mainloop is bound by shared-memory bandwidth
// (TMA fill + MMA operand fetch ~190 B/clk wanted vs 128 B/clk), not by the tensor pipe and not by L2 (a B-multicast
// variant that only cut the L2 reads gained 1%).  Protocol: each CTA's TMA producer fills its own stage, both counting their
// bytes on the LEADER's full barrier (cp.async.bulk.tensor .cta_group::2); the leader's tcgen05.commit is multicast
// onto both CTAs' empty / accumulator-full barriers; both CTAs' epilogue warps drain their own TMEM lanes and release
// the accumulator on the leader's barrier.
//
// Scheduling is "data-parallel + stream-K": whole waves of tiles are dealt round-robin to the persistent CTAs; the
// remainder tiles (all tiles, when there are fewer tiles than SMs) form a stream-K region whose (tile, k-block)
// iteration space is cut into one contiguous range per CTA, done BEFORE the CTA's whole tiles -- every SM gets the same
// number of k-blocks (no wave quantisation, no split-K workspace pass).  A tile whose k-range is shared is finished by
// the CTA that owns its FIRST k-block; the other CTAs reach their share of it first thing, park their raw fp32
// accumulators in a per-CTA workspace slot and raise a flag.  The owner adds the slots in CTA order, so results are
// run-to-run deterministic.
//
// Replaces the ATen calls behind nn.Conv2d / nn.Linear / MHA projections on the reference hot path
// (src/models/unet.py:80,90,96,160-187,325-399); SURVEY.md §2.1 K1-K3, K6-K9.
#include <cstdlib>
#include "gemm_desc.h"
#include <cuda.h>
#include <string.h>

namespace umma {

constexpr int BLOCK_M = 128;
constexpr int BLOCK_K = 64;  // 64 bf16 = 128 B = one SWIZZLE_128B row
constexpr int UMMA_K = 16;
constexpr int kEpiWarps = 8;       // two warps per TMEM lane quarter (16 were measured: no faster, and they spill)
constexpr int kNumThreads = 64 + 32 * kEpiWarps;   // warp 0 TMA, warp 1 MMA, warps 2.. epilogue
constexpr uint32_t A_BYTES = BLOCK_M * BLOCK_K * 2;

struct alignas(64) KernelParams {
  CUtensorMap tm_a;
  CUtensorMap tm_b;
  int M, N;
  int num_kb;
  int num_m_tiles, num_n_tiles;
  int num_tiles;             // num_m_tiles * num_n_tiles
  int sk_tiles, sk_ctas;     // stream-K region: tiles [0, sk_tiles) shared by CTAs [0, sk_ctas)
  int sk_split;              // > 0: aligned mode, every stream-K tile is cut into sk_split equal k-ranges (one per CTA)
  long long sk_units;        // sk_tiles * num_kb
  float* sk_slots;           // [gridDim.x][kMT*128*kBlockN] fp32 partial accumulators of shared tiles
  int* sk_flags;             // [gridDim.x] 1 = slot holds a partial that has not been consumed yet
  int debug;                 // profiling aid (psg_umma_debug): 1 = skip the epilogue body, 2 = epilogue math without global I/O
  // mode 0, A im2col
  int a_im2col, cblks, ksize, conv_stride, pad, flip, P, Q;
  // mode 1, B im2col (wgrad): columns are (tap, cin)
  int b_im2col, cin, tiles_per_tap;
  // mode 2, B = conv weight [Cout][tap*Cin] read as B(n = cin, k = (tap, cout)): k-block -> (tap, 64-row block of Cout)
  int b_tapped, b_cblks, b_tap_cols;
  PsgEpilogue epi;
};

__device__ int g_timeout_flag = 0;

// ---------------------------------------------------------------------------------------------
// PTX wrappers
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

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
// Bounded wait: a protocol bug must never hang the GPU (a hung box is a strike); on timeout raise a flag
// the host can read (psg_umma_timeout_flag) and fall through.
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
  for (uint32_t i = 0; i < (1u << 24); ++i)
    if (mbar_try_wait(bar, parity)) return;
  atomicExch(&g_timeout_flag, 1);
}

// Polite wait for the epilogue warps: while the mainloop runs they have nothing to do, and eight warps spinning on
// try_wait took issue slots from the single TMA and MMA threads that share their schedulers (ncu: 70% of all samples sat in
// this loop; the per-k-block cost of the producer, not the tensor pipe, bounded the 128-row tiles).
__device__ __forceinline__ void mbar_wait_backoff(uint32_t bar, uint32_t parity) {
  for (uint32_t i = 0; i < (1u << 22); ++i) {
    if (mbar_try_wait(bar, parity)) return;
    __nanosleep(128);
  }
  atomicExch(&g_timeout_flag, 1);
}

__device__ __forceinline__ void tma_load_2d(uint32_t dst, const CUtensorMap* tm, uint32_t bar, int c0, int c1) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];" ::"r"(dst),
      "l"(reinterpret_cast<uint64_t>(tm)), "r"(bar), "r"(c0), "r"(c1)
      : "memory");
}
__device__ __forceinline__ void tma_load_im2col_4d(uint32_t dst, const CUtensorMap* tm, uint32_t bar, int c, int w, int h,
                                                   int n, uint16_t off_w, uint16_t off_h) {
  asm volatile(
      "cp.async.bulk.tensor.4d.shared::cluster.global.im2col.mbarrier::complete_tx::bytes"
      " [%0], [%1, {%3, %4, %5, %6}], [%2], {%7, %8};" ::"r"(dst),
      "l"(reinterpret_cast<uint64_t>(tm)), "r"(bar), "r"(c), "r"(w), "r"(h), "r"(n), "h"(off_w), "h"(off_h)
      : "memory");
}
// cta_group::2 forms: the data lands in THIS CTA's shared memory, the bytes are counted on `bar`, a shared::cluster address
// that may name the barrier of the pair's leader CTA.
__device__ __forceinline__ void tma_load_2d_2cta(uint32_t dst, const CUtensorMap* tm, uint32_t bar, int c0, int c1) {
  asm volatile(
      "cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];" ::"r"(dst),
      "l"(reinterpret_cast<uint64_t>(tm)), "r"(bar), "r"(c0), "r"(c1)
      : "memory");
}
__device__ __forceinline__ void tma_load_im2col_4d_2cta(uint32_t dst, const CUtensorMap* tm, uint32_t bar, int c, int w, int h, int n,
                                                        uint16_t off_w, uint16_t off_h) {
  asm volatile(
      "cp.async.bulk.tensor.4d.im2col.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes"
      " [%0], [%1, {%3, %4, %5, %6}], [%2], {%7, %8};" ::"r"(dst),
      "l"(reinterpret_cast<uint64_t>(tm)), "r"(bar), "r"(c), "r"(w), "r"(h), "r"(n), "h"(off_w), "h"(off_h)
      : "memory");
}
template <int kCl>
__device__ __forceinline__ void tma_2d(uint32_t dst, const CUtensorMap* tm, uint32_t bar, int c0, int c1) {
  if (kCl == 1) tma_load_2d(dst, tm, bar, c0, c1);
  else tma_load_2d_2cta(dst, tm, bar, c0, c1);
}
template <int kCl>
__device__ __forceinline__ void tma_im2col(uint32_t dst, const CUtensorMap* tm, uint32_t bar, int c, int w, int h, int n, uint16_t off_w,
                                           uint16_t off_h) {
  if (kCl == 1) tma_load_im2col_4d(dst, tm, bar, c, w, h, n, off_w, off_h);
  else tma_load_im2col_4d_2cta(dst, tm, bar, c, w, h, n, off_w, off_h);
}
__device__ __forceinline__ void cluster_sync_all() {
  asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}
__device__ __forceinline__ uint32_t cluster_ctarank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}
__device__ __forceinline__ void prefetch_tmap(const CUtensorMap* tm) {
  asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(tm)) : "memory");
}

__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

template <int kCl>
__device__ __forceinline__ void tmem_alloc(uint32_t dst_smem, uint32_t ncols) {
  if (kCl == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(dst_smem), "r"(ncols) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  } else {      // one warp of EACH CTA of the pair executes this; both get the same column base
    asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(dst_smem), "r"(ncols) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
  }
}
template <int kCl>
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr, uint32_t ncols) {
  if (kCl == 1) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
  else          asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}
__device__ __forceinline__ void umma_bf16_2cta(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accum) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t}" ::"r"(tmem_d),
      "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accum)
      : "memory");
}
// the same barrier in the leader CTA (rank 0) of the pair, as a shared::cluster address
__device__ __forceinline__ uint32_t leader_addr(uint32_t local_bar) {
  uint32_t r;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(local_bar), "r"(0));
  return r;
}
// relaxed on purpose: `.release.cluster` compiles to MEMBAR.ALL.GPU + ERRBAR + CGAERRBAR (~1 us), once per k-block that
// serialised the whole pipeline (measured: 2x slower than one CTA).  What the arrive publishes was written by the async
// proxy (TMA complete_tx already observed) or lives in TMEM (ordered by tcgen05.fence), not by this thread's stores.
__device__ __forceinline__ void mbar_arrive_remote(uint32_t cluster_bar) {
  asm volatile("mbarrier.arrive.relaxed.cluster.shared::cluster.b64 _, [%0];" ::"r"(cluster_bar) : "memory");
}
__device__ __forceinline__ void umma_bf16(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accum) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}" ::"r"(tmem_d),
      "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accum)
      : "memory");
}
__device__ __forceinline__ void umma_commit(uint32_t bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}
// arrives on the barrier at the same shared-memory offset in every CTA of `mask` once the MMAs issued so far have completed
__device__ __forceinline__ void umma_commit_mc(uint32_t bar, uint16_t mask) {
  asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(bar), "h"(mask)
               : "memory");
}
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
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}

// Shared-memory matrix descriptors (cute/arch/mma_sm100_desc.hpp bit layout), SWIZZLE_128B.
//  K-major : rows of 128 B, 8-row groups 1024 B apart (SBO); LBO unused (1).
//  MN-major: K-rows of 128 B (64 MN elements), 8-row groups 1024 B apart (SBO), 64-element MN groups `lbo` apart.
__device__ __forceinline__ uint64_t make_desc(uint32_t saddr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
  uint64_t d = 0;
  d |= (uint64_t)((saddr & 0x3FFFF) >> 4);
  d |= (uint64_t)((lbo_bytes >> 4) & 0x3FFF) << 16;
  d |= (uint64_t)((sbo_bytes >> 4) & 0x3FFF) << 32;
  d |= (uint64_t)1 << 46;  // descriptor version (Blackwell)
  d |= (uint64_t)2 << 61;  // SWIZZLE_128B
  return d;
}
__host__ __device__ constexpr uint32_t make_idesc(int n, int a_mn, int b_mn, int m = BLOCK_M) {
  return (1u << 4)                      // D = f32
         | (1u << 7) | (1u << 10)       // A, B = bf16
         | ((uint32_t)a_mn << 15) | ((uint32_t)b_mn << 16) | ((uint32_t)(n >> 3) << 17) | ((uint32_t)(m >> 4) << 24);
}

// ---------------------------------------------------------------------------------------------
// Epilogue for one warp's 32-row x 32-column chunk of the accumulator.
//
// tcgen05.ld hands every thread ONE ROW (32 consecutive columns).  Touching global memory in that shape makes each
// warp-level 16-byte access hit 32 different rows (32 L2 requests of 16 bytes): the epilogue, not the tensor pipe, was
// what bounded the short-K GEMMs.  So every tensor the epilogue reads or writes goes through a per-warp staging tile in
// shared memory (XOR-swizzled, conflict-free both ways): global accesses are issued "piece-major" -- consecutive lanes
// take consecutive 16-byte pieces of a row, 64 or 128 contiguous bytes per row and instruction -- and each thread
// picks up / drops off its own row on the shared-memory side.
// ---------------------------------------------------------------------------------------------
// One staging tile = 32 rows x 64 bytes (32 bf16 columns, or 16 fp32 columns: fp32 tensors move in two halves).
struct WarpTile {
  __device__ static __forceinline__ uint32_t off(int r, int c) { return (uint32_t)(r * 64 + ((c ^ ((r >> 1) & 3)) << 4)); }
  // global tile (row pitch ld_bytes) -> staging; rows >= rows_valid are skipped
  __device__ static __forceinline__ void load(uint8_t* stage, const uint8_t* g, long long ld_bytes, int rows_valid, int lane) {
    uint4 v[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const int pc = lane + 32 * i, r = pc >> 2, c = pc & 3;
      v[i] = (r < rows_valid) ? __ldcg(reinterpret_cast<const uint4*>(g + (long long)r * ld_bytes + c * 16)) : make_uint4(0u, 0u, 0u, 0u);
    }
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const int pc = lane + 32 * i, r = pc >> 2, c = pc & 3;
      *reinterpret_cast<uint4*>(stage + off(r, c)) = v[i];
    }
  }
  __device__ static __forceinline__ void store(const uint8_t* stage, uint8_t* g, long long ld_bytes, int rows_valid, int lane) {
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const int pc = lane + 32 * i, r = pc >> 2, c = pc & 3;
      if (r < rows_valid) __stcg(reinterpret_cast<uint4*>(g + (long long)r * ld_bytes + c * 16), *reinterpret_cast<const uint4*>(stage + off(r, c)));
    }
  }
};

// whole-warp: fetch the [32 x 32] tile at (m_base, n0) of a row-pitched tensor and hand every thread its row
__device__ __forceinline__ void tile_in(uint8_t* stage, const void* base, long long ld, int dtype, long long m_base, long long n0,
                                        int rows_valid, int lane, float (&a)[32]) {
  if (dtype == PSG_DTYPE_BF16) {
    __syncwarp();
    WarpTile::load(stage, reinterpret_cast<const uint8_t*>(base) + (m_base * ld + n0) * 2, ld * 2, rows_valid, lane);
    __syncwarp();
#pragma unroll
    for (int c = 0; c < 4; ++c) {
      const uint4 r = *reinterpret_cast<const uint4*>(stage + WarpTile::off(lane, c));
      const __nv_bfloat162* h = reinterpret_cast<const __nv_bfloat162*>(&r);
#pragma unroll
      for (int i = 0; i < 4; ++i) { const float2 f = __bfloat1622float2(h[i]); a[8 * c + 2 * i] = f.x; a[8 * c + 2 * i + 1] = f.y; }
    }
  } else {
#pragma unroll
    for (int half = 0; half < 2; ++half) {
      __syncwarp();
      WarpTile::load(stage, reinterpret_cast<const uint8_t*>(base) + (m_base * ld + n0 + 16 * half) * 4, ld * 4, rows_valid, lane);
      __syncwarp();
#pragma unroll
      for (int c = 0; c < 4; ++c) {
        const float4 r = *reinterpret_cast<const float4*>(stage + WarpTile::off(lane, c));
        a[16 * half + 4 * c] = r.x; a[16 * half + 4 * c + 1] = r.y; a[16 * half + 4 * c + 2] = r.z; a[16 * half + 4 * c + 3] = r.w;
      }
    }
  }
}
__device__ __forceinline__ void tile_out(uint8_t* stage, void* base, long long ld, int dtype, long long m_base, long long n0,
                                         int rows_valid, int lane, const float (&v)[32]) {
  if (dtype == PSG_DTYPE_BF16) {
    __syncwarp();
#pragma unroll
    for (int c = 0; c < 4; ++c) {
      uint4 r;
      __nv_bfloat162* h = reinterpret_cast<__nv_bfloat162*>(&r);
#pragma unroll
      for (int i = 0; i < 4; ++i) h[i] = __floats2bfloat162_rn(v[8 * c + 2 * i], v[8 * c + 2 * i + 1]);
      *reinterpret_cast<uint4*>(stage + WarpTile::off(lane, c)) = r;
    }
    __syncwarp();
    WarpTile::store(stage, reinterpret_cast<uint8_t*>(base) + (m_base * ld + n0) * 2, ld * 2, rows_valid, lane);
  } else {
#pragma unroll
    for (int half = 0; half < 2; ++half) {
      __syncwarp();
#pragma unroll
      for (int c = 0; c < 4; ++c)
        *reinterpret_cast<float4*>(stage + WarpTile::off(lane, c)) =
            make_float4(v[16 * half + 4 * c], v[16 * half + 4 * c + 1], v[16 * half + 4 * c + 2], v[16 * half + 4 * c + 3]);
      __syncwarp();
      WarpTile::store(stage, reinterpret_cast<uint8_t*>(base) + (m_base * ld + n0 + 16 * half) * 4, ld * 4, rows_valid, lane);
    }
  }
}

// L2 prefetch of this thread's row segment (64 B bf16 / 128 B fp32) of an epilogue input tile: issued before the wait
// for the accumulator, so that the tile_in loads that follow find their lines in L2.
__device__ __forceinline__ void prefetch_row(const void* base, long long ld, int dtype, long long m, long long n0) {
  const uint8_t* p = reinterpret_cast<const uint8_t*>(base) + (m * ld + n0) * (dtype == PSG_DTYPE_BF16 ? 2 : 4);
  asm volatile("prefetch.global.L2 [%0];" ::"l"(p));
  if (dtype != PSG_DTYPE_BF16) asm volatile("prefetch.global.L2 [%0];" ::"l"(p + 64));
}

// Called by the whole warp (rows m_base .. m_base+31, this thread's row m = m_base + lane; rows >= M are dead).
__device__ __forceinline__ void epilogue_chunk(const PsgEpilogue& e, const uint32_t (&acc)[32], long long m_base, int lane,
                                               long long M, long long n0, int nvalid, long long N, uint8_t* stage) {
  const long long m = m_base + lane;
  const int rows_valid = (M - m_base) < 32 ? (int)(M - m_base) : 32;
  // fast path: whole chunk valid and 16B-aligned everywhere (warp-uniform condition)
  const bool vec = (nvalid == 32) && ((n0 & 7) == 0) && ((e.ldc & 7) == 0) && ((e.ldr & 7) == 0) && ((e.ld_aux & 7) == 0) &&
                   ((e.ld_rowbias & 3) == 0) && (((N & 1) == 0) || !e.drop_threshold);
  if (!vec) {
    if (m < M) {
#pragma unroll
      for (int j = 0; j < 32; ++j)       // (fully unrolled on purpose: a runtime index would push acc[] into local memory)
        if (j < nvalid) psg_epilogue_scalar(e, __uint_as_float(acc[j]), m, n0 + j, N);
    }
    return;
  }
  const bool live = m < M;
  float v[32];
#pragma unroll
  for (int j = 0; j < 32; ++j) v[j] = __uint_as_float(acc[j]);
  if (e.bias) {
    const float4* b4 = reinterpret_cast<const float4*>(e.bias + n0);
#pragma unroll
    for (int j = 0; j < 8; ++j) { float4 b = __ldg(b4 + j); v[4*j] += b.x; v[4*j+1] += b.y; v[4*j+2] += b.z; v[4*j+3] += b.w; }
  }
  if (e.rowbias && live) {
    const float4* b4 = reinterpret_cast<const float4*>(e.rowbias + (m / e.rows_per_group) * e.ld_rowbias + n0);
#pragma unroll
    for (int j = 0; j < 8; ++j) { float4 b = __ldg(b4 + j); v[4*j] += b.x; v[4*j+1] += b.y; v[4*j+2] += b.z; v[4*j+3] += b.w; }
  }
  // dropout decisions of this thread's 32 elements, one bit each (one hash per neighbouring pair)
  uint32_t keep = 0xFFFFFFFFu;
  if (e.drop_threshold) {
    keep = 0u;
    const uint64_t pair0 = (uint64_t)(m * N + n0) >> 1;      // N and n0 are even here: (m*N + n0) is even
#pragma unroll
    for (int j = 0; j < 32; j += 2) {
      const uint32_t h = psg_hash32(e.drop_seed, pair0 + (j >> 1));
      keep |= (psg_drop_keep2(h, 0, e.drop_threshold) ? 1u : 0u) << j;
      keep |= (psg_drop_keep2(h, 1, e.drop_threshold) ? 1u : 0u) << (j + 1);
    }
  }
  const bool save_grad = e.aux_out && e.aux_act != PSG_ACT_NONE;   // save act'(pre) * dropmask instead of pre
  if (e.aux_out && !save_grad) tile_out(stage, e.aux_out, e.ld_aux, e.act_dtype, m_base, n0, rows_valid, lane, v);
  if (save_grad) {
    float d[32];
    if (e.act == PSG_ACT_GELU) {
#pragma unroll
      for (int j = 0; j < 32; ++j) {
        float cdf, ex;
        psg_gelu_parts(v[j], cdf, ex);
        d[j] = fmaf(v[j] * 0.3989422804014327f, ex, cdf);
        v[j] *= cdf;
      }
    } else if (e.act == PSG_ACT_SILU) {
#pragma unroll
      for (int j = 0; j < 32; ++j) {
        const float sg = psg_rcp_approx(1.f + psg_ex2_approx(-1.4426950408889634f * v[j]));
        d[j] = sg * fmaf(v[j], 1.f - sg, 1.f);
        v[j] *= sg;
      }
    } else {
#pragma unroll
      for (int j = 0; j < 32; ++j) d[j] = 1.f;
    }
#pragma unroll
    for (int j = 0; j < 32; ++j) d[j] = ((keep >> j) & 1u) ? d[j] * e.drop_scale : 0.f;
    tile_out(stage, e.aux_out, e.ld_aux, e.act_dtype, m_base, n0, rows_valid, lane, d);
  } else if (e.act == PSG_ACT_GELU) {
#pragma unroll
    for (int j = 0; j < 32; ++j) v[j] = psg_gelu_fast(v[j]);
  } else if (e.act == PSG_ACT_SILU) {
#pragma unroll
    for (int j = 0; j < 32; ++j) v[j] = psg_silu_fast(v[j]);
  } else if (e.act == PSG_ACT_RELU) {
#pragma unroll
    for (int j = 0; j < 32; ++j) v[j] = fmaxf(v[j], 0.f);
  }
  if (e.aux_in) {
    float a[32];
    tile_in(stage, e.aux_in, e.ld_aux, e.act_dtype, m_base, n0, rows_valid, lane, a);
    if (e.aux_act == PSG_ACT_MUL) {
#pragma unroll
      for (int j = 0; j < 32; ++j) v[j] *= a[j];
    } else if (e.aux_act == PSG_ACT_GELU) {
#pragma unroll
      for (int j = 0; j < 32; ++j) v[j] *= psg_gelu_grad_fast(a[j]);
    } else if (e.aux_act == PSG_ACT_SILU) {
#pragma unroll
      for (int j = 0; j < 32; ++j) v[j] *= psg_silu_grad_fast(a[j]);
    }
  }
  if (e.drop_threshold) {
#pragma unroll
    for (int j = 0; j < 32; ++j) v[j] = ((keep >> j) & 1u) ? v[j] * e.drop_scale : 0.f;
  }
  if (e.alpha != 1.f) {
#pragma unroll
    for (int j = 0; j < 32; ++j) v[j] *= e.alpha;
  }
  if (e.residual) {
    float a[32];
    tile_in(stage, e.residual, e.ldr, e.act_dtype, m_base, n0, rows_valid, lane, a);
#pragma unroll
    for (int j = 0; j < 32; ++j) v[j] += a[j];
  }
  if (e.out_dtype == PSG_DTYPE_F32 && e.accumulate) {
    float a[32];
    tile_in(stage, e.out, e.ldc, PSG_DTYPE_F32, m_base, n0, rows_valid, lane, a);
#pragma unroll
    for (int j = 0; j < 32; ++j) v[j] += a[j];
  }
  tile_out(stage, e.out, e.ldc, e.out_dtype, m_base, n0, rows_valid, lane, v);
}

__host__ __device__ constexpr uint32_t tmem_cols_for(int n) { return n <= 32 ? 32 : n <= 64 ? 64 : n <= 128 ? 128 : n <= 256 ? 256 : 512; }

// kMT = number of 128-row MMA tiles stacked in one CTA tile (BLOCK_M = 128 * kMT).  The two MMAs of a k-step share
// the B tile in shared memory, which halves the L2->SM bytes per FLOP of the weight operand; measured on B200 the
// engine is bound by operand delivery (bytes per FLOP), not by the tensor pipe, so larger CTA tiles are what moves it.
template <int kBlockN, int kStages, int kMT, int kCl = 1>
struct SmemLayout {
  static constexpr uint32_t A_TILE_BYTES = A_BYTES * kMT;
  static constexpr uint32_t B_BYTES = kBlockN * BLOCK_K * 2 / kCl;      // a CTA of a pair holds half of the B tile
  static constexpr uint32_t STAGE_BYTES = A_TILE_BYTES + B_BYTES;
  static constexpr int kAcc = (2 * kMT * kBlockN <= 512) ? 2 : 1;          // TMEM accumulator stages
  static constexpr uint32_t TMEM_COLS = tmem_cols_for(kAcc * kMT * kBlockN);
  static constexpr uint32_t EPI_OFFSET = STAGE_BYTES * kStages;         // one 2 KiB staging tile per epilogue warp
  static constexpr uint32_t EPI_BYTES = kEpiWarps * 2048;
  static constexpr uint32_t BAR_OFFSET = EPI_OFFSET + EPI_BYTES;
  static constexpr uint32_t TOTAL = BAR_OFFSET + (3 * kStages + 4) * 8 + 16 + 1024;  // + alignment slack
};

__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}

// Work decomposition.  Stream-K region: unit u = tile * num_kb + kb over tiles [0, sk_tiles); CTA c < sk_ctas owns units
// [sk_begin(c), sk_begin(c+1)).  Data-parallel region: CTA c then takes tiles sk_tiles + c, sk_tiles + c + gridDim.x, ...
// Aligned mode (sk_split > 0): CTA c takes the (c % split)-th k-range of tile c / split, so that the CTAs of all tiles
// walk the SAME k positions at the same time and share every operand slab in L2 (wgrad over activations larger than L2:
// with unaligned ranges each tile re-read its operands from DRAM, 60 GB per step).
__device__ __forceinline__ long long sk_begin(long long sk_units, int c, int sk_ctas, int split, int num_kb) {
  if (c >= sk_ctas) return sk_units;
  if (split > 0) {
    const int tile = c / split, s = c - tile * split;
    return (long long)tile * num_kb + (long long)s * num_kb / split;
  }
  return sk_units * c / sk_ctas;
}

struct Segment {          // a maximal run of one CTA's k-blocks inside one tile
  int tile, kb0, kb1;
};
struct SegmentIter {
  long long cur, end;
  int num_kb, dp_tile, num_tiles, stride;
  __device__ __forceinline__ SegmentIter(const KernelParams& p, int c, int ctas)
      : cur(sk_begin(p.sk_units, c, p.sk_ctas, p.sk_split, p.num_kb)), end(sk_begin(p.sk_units, c + 1, p.sk_ctas, p.sk_split, p.num_kb)),
        num_kb(p.num_kb), dp_tile(p.sk_tiles + c), num_tiles(p.num_tiles), stride(ctas) {}
  __device__ __forceinline__ bool next(Segment& s) {
    if (cur < end) {
      s.tile = (int)(cur / num_kb);
      s.kb0 = (int)(cur - (long long)s.tile * num_kb);
      const long long left = end - cur;
      s.kb1 = (left < (long long)(num_kb - s.kb0)) ? s.kb0 + (int)left : num_kb;
      cur += s.kb1 - s.kb0;
      return true;
    }
    if (dp_tile < num_tiles) {
      s.tile = dp_tile;
      s.kb0 = 0;
      s.kb1 = num_kb;
      dp_tile += stride;
      return true;
    }
    return false;
  }
};
#define PSG_SEGMENTS(p) SegmentIter segs((p), vcta, vctas)

__device__ __forceinline__ void epi_bar_sync() { asm volatile("bar.sync 1, %0;" ::"n"(32 * kEpiWarps) : "memory"); }   // the epilogue warps

// Bounded spin on a stream-K flag (see mbar_wait: a protocol bug must not hang the box).
__device__ __forceinline__ void wait_flag(const int* flag) {
  for (uint32_t i = 0; i < (1u << 22); ++i) {
    int v;
    asm volatile("ld.acquire.gpu.global.s32 %0, [%1];" : "=r"(v) : "l"(flag) : "memory");
    if (v != 0) return;
    __nanosleep(64);
  }
  atomicExch(&g_timeout_flag, 1);
}

// ---------------------------------------------------------------------------------------------
// the kernel: persistent CTAs, one contiguous stream-K range of (tile, k-block) units each
// ---------------------------------------------------------------------------------------------
// Sub-tiles (128 rows per CTA, 128 * kCl per scheduling unit) of row tile `tile_m` that hold at least one row < M.  With kMT = 2 the
// last row tile of e.g. M = 640 (pairs: 512-row tiles) or 1280 keeps only its lower half: the upper one is neither loaded, multiplied
// nor drained (all three roles derive the same count).
template <int kMT, int kCl>
__device__ __forceinline__ int live_subtiles(int tile_m, int M) {
  if (kMT == 1) return 1;
  return (long long)tile_m * (BLOCK_M * kMT * kCl) + BLOCK_M * kCl >= M ? 1 : kMT;
}

template <int kBlockN, int kStages, int kMode, int kMT, int kCl>
__global__ void __launch_bounds__(kNumThreads, 1) umma_gemm_kernel(const __grid_constant__ KernelParams p) {
  using L = SmemLayout<kBlockN, kStages, kMT, kCl>;
  constexpr int kAcc = L::kAcc;
  constexpr int BM = BLOCK_M * kMT;          // rows of this CTA's share of the tile
  static_assert(kBlockN % 32 == 0 && kBlockN <= 256, "BLOCK_N must be a multiple of 32 (epilogue chunk) and <= 256");
  static_assert(kMode == 0 || kBlockN % 64 == 0, "MN-major B tiles are built from 64-wide TMA boxes");
  constexpr bool kAMn = (kMode == 1), kBMn = (kMode >= 1);
  static_assert(kCl == 1 || (kCl == 2 && kBlockN % 128 == 0), "CTA pairs split the B tile in two halves of whole 64-wide boxes");
  const uint32_t cta_rank = (kCl == 2) ? cluster_ctarank() : 0u;     // which half of B / which 128-row halves of the tile are this CTA's
  const int vcta = blockIdx.x / kCl, vctas = gridDim.x / kCl;        // scheduling unit: a CTA, or a CTA pair
  static_assert(kMT * kBlockN <= 512, "accumulators exceed TMEM");
  extern __shared__ uint8_t smem_raw[];
  const uint32_t smem = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const uint32_t bar_full = smem + L::BAR_OFFSET;
  const uint32_t bar_empty = bar_full + 8 * kStages;
  const uint32_t bar_pfull = bar_empty + 8 * kStages;     // [kStages] pair only, used in the leader: the peer's stage is full
  const uint32_t bar_tfull = bar_pfull + 8 * kStages;     // [kAcc] accumulator ready for the epilogue
  const uint32_t bar_tempty = bar_tfull + 16;             // [kAcc] accumulator drained by the epilogue
  const uint32_t tmem_slot = bar_tempty + 16;
  uint32_t* tmem_slot_ptr = reinterpret_cast<uint32_t*>(smem_raw + (tmem_slot - smem_u32(smem_raw)));

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;

  if (warp == 0 && lane == 0) {
    prefetch_tmap(&p.tm_a);
    prefetch_tmap(&p.tm_b);
    for (int s = 0; s < kStages; ++s) { mbar_init(bar_full + 8 * s, 1); mbar_init(bar_empty + 8 * s, 1); mbar_init(bar_pfull + 8 * s, 1); }
    for (int a = 0; a < kAcc; ++a) { mbar_init(bar_tfull + 8 * a, 1); mbar_init(bar_tempty + 8 * a, kCl * kEpiWarps); }   // both CTAs' epilogues
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 1) tmem_alloc<kCl>(tmem_slot, L::TMEM_COLS);
  tc_fence_before();
  __syncthreads();
  if (kCl == 2) cluster_sync_all();        // the peer's barriers are initialised before anything arrives on them
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot_ptr;

  if (warp == 0) {
    if (lane == 0) {
      // ===== TMA producer =====
      uint32_t it = 0;   // global k-block counter: smem ring position and phase carry across segments
      PSG_SEGMENTS(p);
      Segment sg;
      while (segs.next(sg)) {
        const int tile = sg.tile;
        const int tile_n = tile % p.num_n_tiles, tile_m = tile / p.num_n_tiles;
        // rows of this CTA's t-th 128-row sub-tile: m0 + t * kRowStep (alone: consecutive; in a pair the t-th MMA spans 256 rows)
        constexpr int kRowStep = BLOCK_M * kCl;
        const int m0 = tile_m * (BM * kCl) + (int)cta_rank * BLOCK_M;
        const int kb0 = sg.kb0, kb1 = sg.kb1;
        const int live = live_subtiles<kMT, kCl>(tile_m, p.M);      // 1: the upper 128-row sub-tile(s) are past M -- not loaded / issued
        const uint32_t stage_tx = kCl * (L::STAGE_BYTES - (uint32_t)(kMT - live) * A_BYTES);
        int w0[kMT], h0[kMT], img0[kMT];
        if (!kAMn && p.a_im2col) {
          const int pq = p.P * p.Q;
#pragma unroll
          for (int t = 0; t < kMT; ++t) {
            const int mm = m0 + t * kRowStep;
            img0[t] = mm / pq;
            const int rem = mm - img0[t] * pq;
            const int pp = rem / p.Q, qq = rem - pp * p.Q;
            w0[t] = qq * p.conv_stride - p.pad;
            h0[t] = pp * p.conv_stride - p.pad;
          }
        }
        for (int kb = kb0; kb < kb1; ++kb, ++it) {
          const int s = it % kStages;
          const uint32_t ph = (it / kStages) & 1;
          mbar_wait(bar_empty + 8 * s, ph ^ 1);
          // in a pair both CTAs' loads are counted on the LEADER's full barrier (the leader issues the MMAs)
          const uint32_t full = (kCl == 2) ? leader_addr(bar_full + 8 * s) : bar_full + 8 * s;
          if (kCl == 1 || cta_rank == 0) mbar_expect_tx(bar_full + 8 * s, stage_tx);
          const uint32_t sa = smem + s * L::STAGE_BYTES;
          const uint32_t sb = sa + L::A_TILE_BYTES;
          if (!kAMn) {
            if (p.a_im2col) {
              const int tap = kb / p.cblks, cb = kb - tap * p.cblks;
              int r = tap / p.ksize, ss = tap - r * p.ksize;
              if (p.flip) { r = p.ksize - 1 - r; ss = p.ksize - 1 - ss; }
#pragma unroll
              for (int t = 0; t < kMT; ++t)
                if (t < live) tma_im2col<kCl>(sa + t * A_BYTES, &p.tm_a, full, cb * BLOCK_K, w0[t], h0[t], img0[t], (uint16_t)ss, (uint16_t)r);
            } else {
#pragma unroll
              for (int t = 0; t < kMT; ++t)
                if (t < live) tma_2d<kCl>(sa + t * A_BYTES, &p.tm_a, full, kb * BLOCK_K, m0 + t * kRowStep);
            }
          } else {
#pragma unroll
            for (int j = 0; j < BM / 64; ++j)
              if ((j >> 1) < live) tma_2d<kCl>(sa + j * 8192, &p.tm_a, full, m0 + (j >> 1) * kRowStep + 64 * (j & 1), kb * BLOCK_K);
          }
          // B: alone, the whole tile; in a pair, this CTA's half of it (the MMA reads the other half from the peer's smem)
          constexpr int kMyBoxes = kBlockN / 64 / kCl;
          const int jb0 = (int)cta_rank * kMyBoxes;
          if (!kBMn) {
            tma_2d<kCl>(sb, &p.tm_b, full, kb * BLOCK_K, tile_n * kBlockN + (int)cta_rank * (kBlockN / kCl));
          } else if (p.b_im2col) {
            const int pix = kb * BLOCK_K;
            const int pq = p.P * p.Q;
            const int img = pix / pq;
            const int rem = pix - img * pq;
            const int pp = rem / p.Q, qq = rem - pp * p.Q;
            // columns are the flat (tap, cin) index; cin % 64 == 0, so every 64-wide box lies inside one tap
#pragma unroll
            for (int jj = 0; jj < kMyBoxes; ++jj) {
              const int col = tile_n * kBlockN + 64 * (jb0 + jj);
              int tap = col / p.cin;
              int c0 = col - tap * p.cin;
              if (tap >= p.ksize * p.ksize) { tap = 0; c0 = p.cin; }     // past the last tap: channel OOB -> zero fill
              const int r = tap / p.ksize, ss = tap - r * p.ksize;
              tma_im2col<kCl>(sb + jj * 8192, &p.tm_b, full, c0, qq * p.conv_stride - p.pad, pp * p.conv_stride - p.pad, img,
                                 (uint16_t)ss, (uint16_t)r);
            }
          } else {
            int row = kb * BLOCK_K, col0 = tile_n * kBlockN;
            if (p.b_tapped) {          // k-block = (tap, 64 rows of Cout); the tap selects a column band of the weight matrix
              const int tap = kb / p.b_cblks;
              row = (kb - tap * p.b_cblks) * BLOCK_K;
              col0 += tap * p.b_tap_cols;
            }
#pragma unroll
            for (int jj = 0; jj < kMyBoxes; ++jj) tma_2d<kCl>(sb + jj * 8192, &p.tm_b, full, col0 + 64 * (jb0 + jj), row);
          }
        }
      }
    }
    __syncwarp();
  } else if (warp == 1) {
    if (lane == 0) {
      if (kCl == 2 && cta_rank != 0) {
        // ===== peer of a pair: nothing to issue (its TMA loads are counted on the leader's barriers) =====
      } else {
      // ===== MMA issuer (alone, or leader of a pair: one instruction drives both CTAs' tensor cores) =====
      // kNSplit > 1 issues the tile as independent column halves per k-step (experiment: does a single dependent
      // accumulator chain leave a bubble in the tensor pipe? it does not).
      constexpr int kNSplit = 1;   // (2 was measured: 8% slower -- the smaller MMAs cost more than the dependency they remove)
      constexpr int kMmaN = kBlockN / kNSplit;
      constexpr uint32_t idesc = make_idesc(kMmaN, kAMn ? 1 : 0, kBMn ? 1 : 0, BLOCK_M * kCl);
      uint32_t it = 0;
      int wi = 0;   // local segment counter -> accumulator stage and phase
      PSG_SEGMENTS(p);
      Segment sg;
      for (; segs.next(sg); ++wi) {
        const int kb0 = sg.kb0, kb1 = sg.kb1;
        const int live = live_subtiles<kMT, kCl>(sg.tile / p.num_n_tiles, p.M);
        const int acc = wi % kAcc;
        const uint32_t aph = (wi / kAcc) & 1;
        mbar_wait(bar_tempty + 8 * acc, aph ^ 1);       // the epilogue (of both CTAs) has drained this accumulator
        tc_fence_after();
        const uint32_t tmem_acc = tmem_base + acc * (kMT * kBlockN);
        for (int kb = kb0; kb < kb1; ++kb, ++it) {
          const int s = it % kStages;
          const uint32_t ph = (it / kStages) & 1;
          mbar_wait(bar_full + 8 * s, ph);
          tc_fence_after();
          const uint32_t sa = smem + s * L::STAGE_BYTES;
          const uint32_t sb = sa + L::A_TILE_BYTES;
#pragma unroll
          for (int k = 0; k < BLOCK_K / UMMA_K; ++k) {
#pragma unroll
            for (int h = 0; h < kNSplit; ++h) {
              // column half h of the B tile: rows h*kMmaN.. (K-major) / boxes h*kMmaN/64.. (MN-major) = h * B_BYTES / kNSplit bytes
              const uint32_t sbh = sb + h * (L::B_BYTES / kNSplit);
              uint64_t bd;
              if (!kBMn) bd = make_desc(sbh + k * (UMMA_K * 2), 16, 1024);
              else       bd = make_desc(sbh + k * (UMMA_K * 128), 8192, 1024);
#pragma unroll
              for (int t = 0; t < kMT; ++t) {
                if (t >= live) break;
                uint64_t ad;
                if (!kAMn) ad = make_desc(sa + t * A_BYTES + k * (UMMA_K * 2), 16, 1024);
                else       ad = make_desc(sa + t * A_BYTES + k * (UMMA_K * 128), 8192, 1024);
                if (kCl == 1) umma_bf16(tmem_acc + t * kBlockN + h * kMmaN, ad, bd, idesc, (kb > kb0 || k > 0) ? 1u : 0u);
                else          umma_bf16_2cta(tmem_acc + t * kBlockN + h * kMmaN, ad, bd, idesc, (kb > kb0 || k > 0) ? 1u : 0u);
              }
            }
          }
          if (kCl == 1) umma_commit(bar_empty + 8 * s);   // frees the smem slot once these MMAs have read it
          else umma_commit_mc(bar_empty + 8 * s, 3);      // ... in both CTAs
        }
        if (kCl == 1) umma_commit(bar_tfull + 8 * acc);   // accumulator complete
        else umma_commit_mc(bar_tfull + 8 * acc, 3);      // ... for both CTAs' epilogue warps
      }
      }
    }
    __syncwarp();
  } else {
    // ===== epilogue warps =====
    const int q = warp & 3;                  // TMEM lane quarter this warp may access
    const int cslot = (warp - 2) >> 2;       // the warps of a quarter take every (kEpiWarps/4)-th 32-column chunk
    const int row = q * 32 + lane;           // row of the 128-row MMA tile this thread drains
    uint8_t* stage = smem_raw + (smem - smem_u32(smem_raw)) + L::EPI_OFFSET + (warp - 2) * 2048;   // this warp's staging tile
    constexpr int kChunks = kBlockN / 32;
    constexpr long long kSlotFloats = (long long)kMT * BLOCK_M * kBlockN;
    int wi = 0;
    PSG_SEGMENTS(p);
    Segment sg;
    for (; segs.next(sg); ++wi) {
      const int tile = sg.tile;
      const int tile_n = tile % p.num_n_tiles, tile_m = tile / p.num_n_tiles;
      const long long tile_row0 = (long long)tile_m * (BM * kCl) + (long long)cta_rank * BLOCK_M;   // + t * BLOCK_M * kCl per sub-tile
      const int acc = wi % kAcc;
      const uint32_t aph = (wi / kAcc) & 1;
      const int live = live_subtiles<kMT, kCl>(tile_m, p.M);
      const bool contributor = sg.kb0 > 0;                      // someone else owns this tile: park the partial sums
      const bool shared_owner = sg.kb0 == 0 && sg.kb1 < p.num_kb;
      // scheduling units vcta+1 .. last_contrib hold the rest of this tile (their ranges start inside it); in a pair each
      // CTA exchanges partials with the CTAs of the same rank (slot / flag index = unit * kCl + rank)
      int last_contrib = vcta;
      if (shared_owner) {
        const long long tile_end = (long long)(tile + 1) * p.num_kb;
        while (last_contrib + 1 < p.sk_ctas && sk_begin(p.sk_units, last_contrib + 1, p.sk_ctas, p.sk_split, p.num_kb) < tile_end) ++last_contrib;
        if (warp == 2 && lane == 0)
          for (int c = vcta + 1; c <= last_contrib; ++c) wait_flag(p.sk_flags + c * kCl + cta_rank);
        __syncwarp();
        epi_bar_sync();
      }
      const long long col_base = (long long)tile_n * kBlockN;
      const int col_limit = p.N - tile_n * kBlockN;
      if (!contributor && (p.epi.residual || p.epi.aux_in || p.epi.accumulate)) {
        // while the MMAs of this tile are still running: pull the epilogue's input tiles towards L2
#pragma unroll 1
        for (int t = 0; t < kMT; ++t) {
          const long long m = tile_row0 + (long long)t * (BLOCK_M * kCl) + row;
          if (m >= p.M) continue;
#pragma unroll 1
          for (int ch = cslot; ch < kChunks; ch += kEpiWarps / 4) {
            if (col_limit - ch * 32 <= 0) break;
            const long long n0 = col_base + ch * 32;
            if (p.epi.residual) prefetch_row(p.epi.residual, p.epi.ldr, p.epi.act_dtype, m, n0);
            if (p.epi.aux_in) prefetch_row(p.epi.aux_in, p.epi.ld_aux, p.epi.act_dtype, m, n0);
            if (p.epi.accumulate && p.epi.out_dtype == PSG_DTYPE_F32) prefetch_row(p.epi.out, p.epi.ldc, PSG_DTYPE_F32, m, n0);
          }
        }
      }
      mbar_wait_backoff(bar_tfull + 8 * acc, aph);
      tc_fence_after();
      const uint32_t tmem_acc = tmem_base + acc * (kMT * kBlockN);
#pragma unroll 1
      for (int t = 0; t < live; ++t) {
        const long long m_base = tile_row0 + (long long)t * (BLOCK_M * kCl) + q * 32;
#pragma unroll 1
        for (int ch = cslot; ch < kChunks; ch += kEpiWarps / 4) {
          uint32_t accv[32];
          tmem_ld32(tmem_acc + t * kBlockN + ((uint32_t)(q * 32) << 16) + ch * 32, accv);
          // slot layout: [chunk][float4 index j][row] -> a warp's 16-byte accesses are contiguous
          const long long chunk_off = (long long)(t * kChunks + ch) * (BLOCK_M * 32) + row * 4;
          if (contributor) {
            float* dst = p.sk_slots + (long long)blockIdx.x * kSlotFloats + chunk_off;
#pragma unroll
            for (int j = 0; j < 8; ++j)
              __stcg(reinterpret_cast<float4*>(dst + j * (BLOCK_M * 4)),
                     make_float4(__uint_as_float(accv[4 * j]), __uint_as_float(accv[4 * j + 1]), __uint_as_float(accv[4 * j + 2]),
                                 __uint_as_float(accv[4 * j + 3])));
            continue;
          }
          for (int c = vcta + 1; c <= last_contrib; ++c) {       // fixed CTA order: deterministic sums
            const float* src = p.sk_slots + (long long)(c * kCl + cta_rank) * kSlotFloats + chunk_off;
#pragma unroll
            for (int j = 0; j < 8; ++j) {
              const float4 v = __ldcg(reinterpret_cast<const float4*>(src + j * (BLOCK_M * 4)));
              accv[4 * j] = __float_as_uint(__uint_as_float(accv[4 * j]) + v.x);
              accv[4 * j + 1] = __float_as_uint(__uint_as_float(accv[4 * j + 1]) + v.y);
              accv[4 * j + 2] = __float_as_uint(__uint_as_float(accv[4 * j + 2]) + v.z);
              accv[4 * j + 3] = __float_as_uint(__uint_as_float(accv[4 * j + 3]) + v.w);
            }
          }
          int nvalid = col_limit - ch * 32;
          nvalid = nvalid > 32 ? 32 : nvalid;
          if (p.debug == 1) continue;
          if (m_base < p.M && nvalid > 0) epilogue_chunk(p.epi, accv, m_base, lane, p.M, col_base + ch * 32, nvalid, p.N, stage);
        }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) {
        if (kCl == 2 && cta_rank != 0) mbar_arrive_remote(leader_addr(bar_tempty + 8 * acc));   // the leader issues the MMAs
        else mbar_arrive(bar_tempty + 8 * acc);
      }
      if (contributor) {
        __threadfence();                     // partial sums visible device-wide before the flag
        epi_bar_sync();
        if (warp == 2 && lane == 0) {
          int one = 1;
          asm volatile("st.release.gpu.global.s32 [%0], %1;" ::"l"(p.sk_flags + blockIdx.x), "r"(one) : "memory");
        }
      } else if (shared_owner) {
        epi_bar_sync();                      // every epilogue thread has read the slots: hand them back
        if (warp == 2 && lane == 0)
          for (int c = vcta + 1; c <= last_contrib; ++c) {
            int zero = 0;
            asm volatile("st.release.gpu.global.s32 [%0], %1;" ::"l"(p.sk_flags + c * kCl + cta_rank), "r"(zero) : "memory");
          }
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (kCl == 2) cluster_sync_all();        // the peer may still multicast into this CTA's smem / arrive on its barriers
  if (warp == 1) tmem_dealloc<kCl>(tmem_base, L::TMEM_COLS);
}

// ---------------------------------------------------------------------------------------------
// host side: tensor maps + dispatch
// ---------------------------------------------------------------------------------------------
typedef CUresult (*PFN_encodeTiled)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                    const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                    CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
typedef CUresult (*PFN_encodeIm2col)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                     const int*, const int*, cuuint32_t, cuuint32_t, const cuuint32_t*, CUtensorMapInterleave,
                                     CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
static PFN_encodeTiled g_encode_tiled = nullptr;
static PFN_encodeIm2col g_encode_im2col = nullptr;

static int load_driver_fns() {
  if (g_encode_tiled && g_encode_im2col) return PSG_OK;
  cudaDriverEntryPointQueryResult qres;
  void* fn = nullptr;
  cudaError_t e = cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qres);
  if (e != cudaSuccess || qres != cudaDriverEntryPointSuccess || !fn) {
    psg_set_error("cuTensorMapEncodeTiled not available: %s", cudaGetErrorString(e));
    return PSG_ERR_CUDA;
  }
  g_encode_tiled = (PFN_encodeTiled)fn;
  fn = nullptr;
  e = cudaGetDriverEntryPoint("cuTensorMapEncodeIm2col", &fn, cudaEnableDefault, &qres);
  if (e != cudaSuccess || qres != cudaDriverEntryPointSuccess || !fn) {
    psg_set_error("cuTensorMapEncodeIm2col not available: %s", cudaGetErrorString(e));
    return PSG_ERR_CUDA;
  }
  g_encode_im2col = (PFN_encodeIm2col)fn;
  return PSG_OK;
}

// 2D bf16 matrix [rows][ld] with `cols` valid columns; box = {box_cols (=64), box_rows}
static int make_tiled_map(CUtensorMap* tm, const void* ptr, long long rows, long long cols, long long ld, int box_cols,
                          int box_rows) {
  cuuint64_t dims[2] = {(cuuint64_t)cols, (cuuint64_t)rows};
  cuuint64_t strides[1] = {(cuuint64_t)ld * 2};
  cuuint32_t box[2] = {(cuuint32_t)box_cols, (cuuint32_t)box_rows};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = g_encode_tiled(tm, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(ptr), dims, strides, box, estr,
                              CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                              CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    psg_set_error("cuTensorMapEncodeTiled failed (%d): rows=%lld cols=%lld ld=%lld box=%dx%d ptr=%p", (int)r, rows, cols, ld,
                  box_cols, box_rows, ptr);
    return PSG_ERR_CUDA;
  }
  return PSG_OK;
}

// NHWC bf16 tensor, pixel pitch ld; im2col window for a ksize x ksize filter with padding `pad`, traversal stride `stride`.
static int make_im2col_map(CUtensorMap* tm, const PsgOperand& o, int channels_per_pixel, int pixels_per_column) {
  cuuint64_t dims[4] = {(cuuint64_t)o.c, (cuuint64_t)o.w, (cuuint64_t)o.h, (cuuint64_t)o.n};
  cuuint64_t strides[3] = {(cuuint64_t)o.ld * 2, (cuuint64_t)o.ld * 2 * o.w, (cuuint64_t)o.ld * 2 * o.w * o.h};
  // bounding box of filter-window base positions: [-pad, dim - 1 + upper], upper = pad - (ksize-1)
  // flip bits 8..15 (when non-zero): pad_hi + 1, the padding of the bottom / right border when it differs from `pad` (the
  // parity-class convolutions of the stride-2 dgrad need windows that run one row / column past the end only)
  const int pad_hi = ((o.flip >> 8) & 0xFF) ? ((o.flip >> 8) & 0xFF) - 1 : o.pad;
  int lower[2] = {-o.pad, -o.pad};
  int upper[2] = {pad_hi - (o.ksize - 1), pad_hi - (o.ksize - 1)};
  cuuint32_t estr[4] = {1, (cuuint32_t)o.stride, (cuuint32_t)o.stride, 1};
  CUresult r = g_encode_im2col(tm, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, const_cast<void*>(o.ptr), dims, strides, lower, upper,
                               (cuuint32_t)channels_per_pixel, (cuuint32_t)pixels_per_column, estr,
                               CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                               CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    psg_set_error("cuTensorMapEncodeIm2col failed (%d): nhwc=%d,%d,%d,%d ld=%lld k=%d s=%d p=%d", (int)r, o.n, o.h, o.w, o.c,
                  o.ld, o.ksize, o.stride, o.pad);
    return PSG_ERR_CUDA;
  }
  return PSG_OK;
}

template <int kBlockN, int kStages, int kMode, int kMT, int kCl>
static int launch(const KernelParams& kp, dim3 grid, cudaStream_t stream) {
  using L = SmemLayout<kBlockN, kStages, kMT, kCl>;
  static_assert(L::TOTAL <= 232448, "shared memory budget exceeded");
  static bool configured = false;
  auto kern = umma_gemm_kernel<kBlockN, kStages, kMode, kMT, kCl>;
  if (!configured) {
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)L::TOTAL);
    if (e != cudaSuccess) { psg_set_error("umma: cudaFuncSetAttribute(smem=%u): %s", L::TOTAL, cudaGetErrorString(e)); return PSG_ERR_CUDA; }
    configured = true;
  }
  if (kCl == 1) {
    kern<<<grid, kNumThreads, L::TOTAL, stream>>>(kp);
  } else {
    cudaLaunchConfig_t cfg;
    memset(&cfg, 0, sizeof(cfg));
    cfg.gridDim = grid;
    cfg.blockDim = dim3(kNumThreads);
    cfg.dynamicSmemBytes = L::TOTAL;
    cfg.stream = stream;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = kCl;
    attr[0].val.clusterDim.y = 1;
    attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    cudaError_t e = cudaLaunchKernelEx(&cfg, kern, kp);
    if (e != cudaSuccess) { psg_set_error("umma: cluster launch failed: %s", cudaGetErrorString(e)); return PSG_ERR_CUDA; }
  }
  PSG_CHECK_LAUNCH("psg_umma_gemm");
  return PSG_OK;
}

// CTA-pair launch: a pair CTA's stage holds only half of B (BN 128: 24 / 40 KB, BN 256: 32 / 48 KB), so it gets more stages.
template <int kBlockN, int kMode>
static int launch_pair(const KernelParams& kp, dim3 grid, cudaStream_t stream, int m_tiles) {
  if constexpr (kBlockN % 128 == 0) {
    if (m_tiles == 1) return launch<kBlockN, (kBlockN == 128 ? 8 : 6), kMode, 1, 2>(kp, grid, stream);
    return launch<kBlockN, (kBlockN == 128 ? 5 : 4), kMode, 2, 2>(kp, grid, stream);
  } else {
    psg_set_error("psg_umma_gemm: CTA pairs need block_n %% 128 == 0");
    return PSG_ERR_UNSUPPORTED;
  }
}

// How many 2-CTA clusters of this kernel family can be resident at once (one CTA per SM; the SMs of a pair share a GPC).
static int max_resident_pairs() {
  static int pairs = -1;
  if (pairs >= 0) return pairs;
  using L = SmemLayout<256, 6, 1, 2>;
  auto kern = umma_gemm_kernel<256, 6, 0, 1, 2>;
  pairs = 0;
  if (cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)L::TOTAL) != cudaSuccess) { cudaGetLastError(); return pairs; }
  cudaLaunchConfig_t cfg;
  memset(&cfg, 0, sizeof(cfg));
  cfg.gridDim = dim3(2 * 80);
  cfg.blockDim = dim3(kNumThreads);
  cfg.dynamicSmemBytes = L::TOTAL;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = 2;
  attr[0].val.clusterDim.y = 1;
  attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  int n = 0;
  if (cudaOccupancyMaxActiveClusters(&n, kern, &cfg) != cudaSuccess) { cudaGetLastError(); return pairs; }
  pairs = n;
  return pairs;
}

}  // namespace umma

extern "C" {

// Stream-K workspace (caller-owned, registered once per process): [kMaxCtas] flags (zeroed by the caller) followed by
// [kMaxCtas] slots of 2*128*256 fp32 partial accumulators.
static constexpr int kMaxCtas = 160;
static constexpr size_t kSlotBytes = (size_t)2 * 128 * 256 * sizeof(float);
// Host-side engine state is kept PER DEVICE (indexed by the calling thread's current device), so two devices driven from one
// process never share a stream-K workspace or an SM reservation.  Within one device the registered workspace serves ONE
// stream at a time PER LANE: the workspace holds kLanes independent copies (flags + slots), and launches that may be in flight at
// the same time on different streams of one device must name different lanes (psg_umma_gemm_lane; the engine's backward pass runs
// its weight-gradient GEMMs on a second stream with lane 1).  g_debug / g_pairs_on are measurement hooks (tools/, tests/),
// process-wide on purpose.
static constexpr int kMaxDevices = 32;
struct DeviceState {
  void* sk_ws = nullptr;
  int reserve_sms = 0;
};
static DeviceState g_dev[kMaxDevices];
static DeviceState& dev_state() {
  int dev = 0;
  cudaGetDevice(&dev);
  return g_dev[(dev >= 0 && dev < kMaxDevices) ? dev : 0];
}
static int g_debug = 0;
static int g_pairs_on = 1;

// CTA pairs (cta_group::2): 0 = never, 1 = where they were measured to pay (default), 2 = wherever the shape allows.
int psg_umma_pairs(int on) { g_pairs_on = on; return PSG_OK; }
// Number of 2-CTA clusters that can be co-resident on this device (0 if cluster launch is unavailable).
int psg_umma_max_pairs() { return umma::max_resident_pairs(); }

// Profiling aid for tools/bench_shapes.py: 1 = drain TMEM but skip the epilogue body (mainloop time alone).
int psg_umma_debug(int flags) { g_debug = flags; return PSG_OK; }
// SMs the persistent grids leave free (for a concurrently running collective's CTAs: parallel.GradSync).  Returns the
// previous value.
int psg_umma_reserve_sms(int n) {
  DeviceState& ds = dev_state();
  const int prev = ds.reserve_sms;
  if (n >= 0) ds.reserve_sms = n;
  return prev;
}

static constexpr int kLanes = 2;
static constexpr size_t kLaneBytes = 1024 + (size_t)kMaxCtas * kSlotBytes;
size_t psg_umma_workspace_bytes() { return kLanes * kLaneBytes; }

int psg_umma_set_workspace(void* ws, size_t bytes) {
  PSG_CHECK_ARG(ws == nullptr || bytes >= psg_umma_workspace_bytes(), "psg_umma_set_workspace: need %zu bytes, got %zu",
                psg_umma_workspace_bytes(), bytes);
  PSG_CHECK_ARG(((uintptr_t)ws % 256) == 0, "psg_umma_set_workspace: workspace must be 256B aligned");
  dev_state().sk_ws = ws;
  return PSG_OK;
}

// Tile-shape heuristic shared by the auto path and psg_umma_plan.
// A plain (K-major, not im2col) TN product with whole 256-row pair tiles and 256-wide column tiles runs as cta_group::2 pairs of
// 128-row CTAs: each CTA fetches half of the B tile (the operand bytes per FLOP of a 256 x 256 tile) and still keeps two TMEM
// accumulator stages, so the epilogue overlaps the next tile's mainloop -- measured 3-8 % faster than the single 128 x 256 CTA
// on the Linear / projection shapes of a batch-256 step and 10-33 % faster than the 256-row single-CTA tile the long-K rule
// would pick (profiles/r02_bench_linear_tile_sweep.txt); M = 4096 is 2 % slower and keeps the single CTA.
static bool tn_pairs(int mode, int a_plain, long long M, int block_n) {
  static const bool enabled = !(getenv("PSG_TN_PAIRS") && atoi(getenv("PSG_TN_PAIRS")) == 0);      // A/B switch
  return enabled && g_pairs_on >= 1 && mode == 0 && a_plain && block_n == 256 && M >= 8192 && M % (2 * umma::BLOCK_M) == 0;
}

static void plan_tiles(int mode, int b_im2col, int a_plain, long long M, long long N, long long K, int* block_n, int* m_tiles) {
  int bn = *block_n;
  if (bn == 0) {
    if (mode == 0) {
      if (N % 256 == 0 || N >= 512) bn = 256;          // OOB-padded 256-wide tiles beat exact 128/160-wide ones (measured)
      else if (N % 160 == 0) bn = 160;
      else bn = N > 128 ? 256 : (N > 64 ? 128 : 64);
    } else {
      bn = N > 128 ? 256 : (N > 64 ? 128 : 64);      // MN-major B: tiles are built from 64-wide boxes
    }
  }
  int mt = *m_tiles;
  if (mt == 0) {
    const long long pad1 = (M + 127) / 128 * 128, pad2 = (M + 255) / 256 * 256;
    mt = 1;
    // 256-row CTA tiles (two MMAs share the B tile) pay off for long reductions with no extra row padding; short-K
    // GEMMs prefer 128-row tiles with double-buffered TMEM (epilogue overlap).  Stream-K balances any tile count.
    // The conv wgrad (NT over im2col pixels) runs ~1.6x faster per MMA row with 256-row tiles, so it takes them even with
    // padded rows; the plain NT products (Linear wgrad) were measured faster with 128-row tiles unless 256 divides M and the
    // reduction is long (K = 4096, the 4x4 level: 38 vs 54 us at 1280 x 1280, profiles/r02_bench_linear_tile_sweep.txt).
    if (M > 128 && ((mode == 1 && b_im2col && pad2 * 2 <= pad1 * 3) || (mode == 1 && !b_im2col && pad2 == pad1 && N >= 1024 && K >= 8192) ||
                    (mode != 1 && pad2 == pad1 && K >= 2048)))
      mt = 2;
    if (tn_pairs(mode, a_plain, M, bn)) mt = 1;
  }
  *block_n = bn;
  *m_tiles = mt;
}

// Reports the tile shape psg_umma_gemm would use for this problem (block_n / m_tiles: in = request or 0, out = choice).
int psg_umma_plan(const PsgGemmDesc* d, int* block_n, int* m_tiles) {
  PSG_CHECK_ARG(d && block_n && m_tiles, "psg_umma_plan: null pointer");
  const int mode = (d->a.mode == PSG_OP_MNMAJOR) ? 1 : ((d->b.mode == PSG_OP_MNMAJOR || d->b.mode == PSG_OP_CONVW_T) ? 2 : 0);
  plan_tiles(mode, d->b.mode == PSG_OP_IM2COL_T, d->a.mode == PSG_OP_KMAJOR, d->M, d->N, d->K, block_n, m_tiles);
  return PSG_OK;
}

// block_n: 0 = auto, else 64/128/160/256 (160 only for K-major B).  m_tiles: 0 = auto, 1 or 2 (CTA tile = 128*m_tiles rows).
int psg_umma_gemm_lane(const PsgGemmDesc* d, int block_n, int m_tiles, int lane, void* stream) {
  using namespace umma;
  PSG_CHECK_ARG(lane >= 0 && lane < kLanes, "psg_umma_gemm: workspace lane %d out of range [0, %d)", lane, kLanes);
  PSG_CHECK_ARG(d != nullptr, "psg_umma_gemm: null desc");
  PSG_CHECK_ARG(d->in_dtype == PSG_DTYPE_BF16, "psg_umma_gemm: operands must be bf16");
  PSG_CHECK_ARG(d->M > 0 && d->N > 0 && d->K > 0, "psg_umma_gemm: empty problem M=%lld N=%lld K=%lld", d->M, d->N, d->K);
  const int am = d->a.mode, bm = d->b.mode;
  int mode;
  if ((am == PSG_OP_KMAJOR || am == PSG_OP_IM2COL) && bm == PSG_OP_KMAJOR) mode = 0;
  else if (am == PSG_OP_MNMAJOR && (bm == PSG_OP_MNMAJOR || bm == PSG_OP_IM2COL_T)) mode = 1;
  else if ((am == PSG_OP_KMAJOR || am == PSG_OP_IM2COL) && (bm == PSG_OP_MNMAJOR || bm == PSG_OP_CONVW_T)) mode = 2;
  else { psg_set_error("psg_umma_gemm: unsupported operand modes a=%d b=%d", am, bm); return PSG_ERR_UNSUPPORTED; }
  PSG_CHECK_ARG(((uintptr_t)d->a.ptr % 16 == 0) && ((uintptr_t)d->b.ptr % 16 == 0), "psg_umma_gemm: operands must be 16B aligned");
  PSG_CHECK_ARG((d->a.ld % 8 == 0) && (d->b.ld % 8 == 0), "psg_umma_gemm: pitches must be multiples of 8 elements");
  PSG_CHECK_ARG(m_tiles >= 0 && m_tiles <= 2, "psg_umma_gemm: m_tiles must be 0, 1 or 2");
  int rc = load_driver_fns();
  if (rc) return rc;

  KernelParams kp;
  memset(&kp, 0, sizeof(kp));
  kp.M = (int)d->M;
  kp.N = (int)d->N;
  kp.epi = d->epi;
  PSG_CHECK_ARG(d->split_k <= 1, "psg_umma_gemm: split-K is gone (stream-K scheduling balances the k-range itself)");
  plan_tiles(mode, bm == PSG_OP_IM2COL_T, am == PSG_OP_KMAJOR, d->M, d->N, d->K, &block_n, &m_tiles);
  // B-multicast pairs: two CTAs on vertically adjacent row tiles share one B tile (each fetches half of it)
  // CTA pairs (cta_group::2).  Measured on this model's shapes (profiles/r01_bench_gemm_modes.txt): a gain of 10-25% for
  // the NT (wgrad) reductions when the 2x taller pair tile does not add much row padding, nothing for TT and the im2col
  // products (conv fprop / dgrad: equal to a single CTA within noise); round 2: the plain TN products take them too (tn_pairs).
  const DeviceState& ds = dev_state();
  const int free_sms = psg_num_sms() - ds.reserve_sms > 8 ? psg_num_sms() - ds.reserve_sms : 8;
  int cl = 1, units = free_sms;
  if (g_pairs_on && block_n % 128 == 0 && d->M > (long long)BLOCK_M * m_tiles) {
    const long long rows1 = (d->M + BLOCK_M * m_tiles - 1) / (BLOCK_M * m_tiles) * (BLOCK_M * m_tiles);
    const long long rows2 = (d->M + 2 * BLOCK_M * m_tiles - 1) / (2 * BLOCK_M * m_tiles) * (2 * BLOCK_M * m_tiles);
    const bool worth = g_pairs_on == 2 || (mode == 1 && rows2 * 4 <= rows1 * 5) ||
                       (m_tiles == 1 && tn_pairs(mode, am == PSG_OP_KMAJOR, d->M, block_n));
    const int pairs = worth ? max_resident_pairs() : 0;
    if (pairs >= 8) { cl = 2; units = pairs < free_sms / 2 ? pairs : free_sms / 2; }
  }
  if (units > kMaxCtas / cl) units = kMaxCtas / cl;

  long long n_tiles = (d->N + block_n - 1) / block_n;
  kp.num_kb = (int)((d->K + BLOCK_K - 1) / BLOCK_K);
  // ---- A operand ----
  if (mode != 1) {
    if (am == PSG_OP_IM2COL) {
      const PsgOperand& a = d->a;
      PSG_CHECK_ARG(a.c % 64 == 0, "psg_umma_gemm: im2col needs C %% 64 == 0 (C=%d)", a.c);
      PSG_CHECK_ARG(d->K == (long long)a.ksize * a.ksize * a.c, "psg_umma_gemm: K != ksize^2*C");
      PSG_CHECK_ARG(d->M == (long long)a.n * a.p * a.q, "psg_umma_gemm: M != n*p*q");
      rc = make_im2col_map(&kp.tm_a, a, 64, BLOCK_M);
      if (rc) return rc;
      kp.a_im2col = 1; kp.cblks = a.c / 64; kp.ksize = a.ksize; kp.conv_stride = a.stride; kp.pad = a.pad; kp.flip = a.flip & 1;
      kp.P = a.p; kp.Q = a.q;
    } else {
      PSG_CHECK_ARG(d->K % 8 == 0, "psg_umma_gemm: K %% 8 != 0");
      rc = make_tiled_map(&kp.tm_a, d->a.ptr, d->M, d->K, d->a.ld, 64, BLOCK_M);
      if (rc) return rc;
    }
  } else {
    rc = make_tiled_map(&kp.tm_a, d->a.ptr, d->K, d->M, d->a.ld, 64, 64);     // A: [K rows][M cols]
    if (rc) return rc;
  }
  // ---- B operand ----
  if (mode == 0) {
    rc = make_tiled_map(&kp.tm_b, d->b.ptr, d->N, d->K, d->b.ld, 64, block_n / cl);
    if (rc) return rc;
  } else {
    PSG_CHECK_ARG(block_n % 64 == 0, "psg_umma_gemm: MN-major B needs block_n %% 64 == 0");
    if (bm == PSG_OP_IM2COL_T) {
      const PsgOperand& b = d->b;
      PSG_CHECK_ARG(d->N == (long long)b.ksize * b.ksize * b.c, "psg_umma_gemm: N != ksize^2*C");
      PSG_CHECK_ARG(d->K == (long long)b.n * b.p * b.q, "psg_umma_gemm: K != n*p*q");
      PSG_CHECK_ARG(b.c % 64 == 0, "psg_umma_gemm: wgrad im2col needs C %% 64 == 0 (C=%d)", b.c);
      rc = make_im2col_map(&kp.tm_b, b, 64, BLOCK_K);
      if (rc) return rc;
      kp.b_im2col = 1; kp.cin = b.c; kp.tiles_per_tap = 0;
      kp.ksize = b.ksize; kp.conv_stride = b.stride; kp.pad = b.pad; kp.P = b.p; kp.Q = b.q;
    } else if (bm == PSG_OP_CONVW_T) {
      // conv weight [Cout = b.n][taps * Cin], Cin = b.c: B(n = cin, k = tap * Cout + cout)
      const PsgOperand& b = d->b;
      const int taps = b.ksize * b.ksize;
      PSG_CHECK_ARG(b.n % 64 == 0, "psg_umma_gemm: transposed conv weight needs Cout %% 64 == 0 (Cout=%d)", b.n);
      PSG_CHECK_ARG(d->N == b.c && d->K == (long long)taps * b.n, "psg_umma_gemm: transposed conv weight: N != Cin or K != taps*Cout");
      PSG_CHECK_ARG(mode == 2 && am == PSG_OP_IM2COL && d->a.c == b.n && d->a.ksize == b.ksize,
                    "psg_umma_gemm: transposed conv weight pairs with an im2col A over Cout channels");
      rc = make_tiled_map(&kp.tm_b, b.ptr, b.n, (long long)taps * b.c, b.ld, 64, 64);
      if (rc) return rc;
      kp.b_tapped = 1; kp.b_cblks = b.n / 64; kp.b_tap_cols = b.c;
    } else {
      rc = make_tiled_map(&kp.tm_b, d->b.ptr, d->K, d->N, d->b.ld, 64, 64);   // B: [K rows][N cols]
      if (rc) return rc;
    }
  }
  const long long m_tiles_n = (d->M + BLOCK_M * m_tiles * cl - 1) / (BLOCK_M * m_tiles * cl);    // row tiles per scheduling unit
  const long long tiles = m_tiles_n * n_tiles;
  PSG_CHECK_ARG(tiles < 2147483647LL / 2, "psg_umma_gemm: too many tiles");
  kp.num_m_tiles = (int)m_tiles_n;
  kp.num_n_tiles = (int)n_tiles;
  kp.num_tiles = (int)tiles;
  kp.debug = g_debug;
  // whole waves of tiles are dealt round-robin; the remainder is the stream-K region, each CTA's share of it being at
  // least max(8, num_kb / 8) k-blocks (at most ~8 CTAs per tile: the owner adds their partials serially)
  const int sms = units;      // scheduling units: CTAs, or CTA pairs
  const long long sk_tiles = tiles % sms;
  long long ctas = tiles >= sms ? sms : 0, sk_ctas = 0;
  // reduction over activations (NT mode) bigger than the L2 can hold next to everything else: keep the k-ranges of all
  // tiles aligned (see sk_begin); all such problems of this model have fewer tiles than SMs
  const bool aligned_split = mode == 1 && tiles < sms && (double)d->K * (double)(d->M + d->N) * 2.0 > 48e6;
  if (sk_tiles > 0) {
    const long long sk_units = sk_tiles * kp.num_kb;
    const long long min_share = kp.num_kb / 8 > 8 ? kp.num_kb / 8 : 8;
    sk_ctas = sk_units / min_share;
    if (sk_ctas < sk_tiles) sk_ctas = sk_tiles;
    if (sk_ctas > sms) sk_ctas = sms;
    if (aligned_split) {
      long long split = sms / sk_tiles;
      if (split > kp.num_kb / 8) split = kp.num_kb / 8;
      if (split < 1) split = 1;
      kp.sk_split = (int)split;
      sk_ctas = sk_tiles * split;
    }
    if (ctas < sk_ctas) ctas = sk_ctas;
    kp.sk_tiles = (int)sk_tiles;
    kp.sk_ctas = (int)sk_ctas;
    kp.sk_units = sk_units;
    if (sk_ctas != sk_tiles) {     // some tile is shared between CTAs
      PSG_CHECK_ARG(ds.sk_ws != nullptr, "psg_umma_gemm: stream-K workspace not registered on this device (psg_umma_set_workspace)");
      char* ws = reinterpret_cast<char*>(ds.sk_ws) + (size_t)lane * kLaneBytes;
      kp.sk_flags = reinterpret_cast<int*>(ws);
      kp.sk_slots = reinterpret_cast<float*>(ws + 1024);
    }
  }
  dim3 grid((unsigned)(ctas * cl));
  cudaStream_t s = (cudaStream_t)stream;

#define PSG_LAUNCH(BN, ST1, ST2, MODE)                                                            \
  if (cl == 2) return launch_pair<BN, MODE>(kp, grid, s, m_tiles);                                \
  if (m_tiles == 1) return launch<BN, ST1, MODE, 1, 1>(kp, grid, s);                              \
  else return launch<BN, ST2, MODE, 2, 1>(kp, grid, s)
  if (mode == 0) {
    switch (block_n) {
      case 64: PSG_LAUNCH(64, 6, 4, 0);
      case 128: PSG_LAUNCH(128, 6, 4, 0);
      case 160: PSG_LAUNCH(160, 5, 3, 0);
      case 256: PSG_LAUNCH(256, 4, 3, 0);
    }
  } else if (mode == 1) {
    switch (block_n) {
      case 64: PSG_LAUNCH(64, 6, 4, 1);
      case 128: PSG_LAUNCH(128, 6, 4, 1);
      case 256: PSG_LAUNCH(256, 4, 3, 1);
    }
  } else {
    switch (block_n) {
      case 64: PSG_LAUNCH(64, 6, 4, 2);
      case 128: PSG_LAUNCH(128, 6, 4, 2);
      case 256: PSG_LAUNCH(256, 4, 3, 2);
    }
  }
#undef PSG_LAUNCH
  psg_set_error("psg_umma_gemm: unsupported block_n=%d for mode %d", block_n, mode);
  return PSG_ERR_UNSUPPORTED;
}

int psg_umma_gemm_ex(const PsgGemmDesc* d, int block_n, int m_tiles, void* stream) { return psg_umma_gemm_lane(d, block_n, m_tiles, 0, stream); }
int psg_umma_gemm(const PsgGemmDesc* d, int block_n, void* stream) { return psg_umma_gemm_lane(d, block_n, 0, 0, stream); }

// Test hook: 1 if any mbarrier wait timed out since the last call (synchronises the device).
int psg_umma_timeout_flag() {
  int v = 0, zero = 0;
  cudaMemcpyFromSymbol(&v, umma::g_timeout_flag, sizeof(int));
  cudaMemcpyToSymbol(umma::g_timeout_flag, &zero, sizeof(int));
  return v;
}

}  // extern "C"
