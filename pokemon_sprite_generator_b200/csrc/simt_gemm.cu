// SIMT (CUDA-core, fp32 accumulate) GEMM / implicit-GEMM convolution engine.
//
// Role: (1) the fp32 parity mode of the U-Net (north_star: "within 1e-4 max-abs in an fp32 mode"), where
// activations and weights stay fp32 end to end; (2) the layers whose shapes do not fit the tcgen05 tiles
// (init_conv Cin=8, final_conv Cout=8, M = batch-sized conditioning linears); (3) general-stride dgrad gather.
// It consumes the same PsgGemmDesc as the tcgen05 engine, so the two can be checked against each other.
//
// Reference ops replaced: nn.Conv2d / nn.Linear forward+backward (src/models/unet.py:29-33,80-96,325,399).
#include "gemm_desc.h"

namespace simt {

constexpr int TM = 64, TN = 64, TK = 16;
constexpr int kThreads = 256;

template <typename T>
struct OperandView {
  const T* ptr;
  int mode;
  long long ld;
  long long rows, K;
  int n, h, w, c, p, q, stride, pad, ksize, flip;
};

template <typename T>
__device__ __forceinline__ float fetch(const OperandView<T>& o, long long i, long long k) {
  if (i >= o.rows || k >= o.K) return 0.f;
  switch (o.mode) {
    case PSG_OP_KMAJOR: return psg_ld(o.ptr + i * o.ld + k);
    case PSG_OP_MNMAJOR: return psg_ld(o.ptr + k * o.ld + i);
    case PSG_OP_IM2COL:
    case PSG_OP_IM2COL_T: {
      long long pix = (o.mode == PSG_OP_IM2COL) ? i : k;
      long long kk = (o.mode == PSG_OP_IM2COL) ? k : i;
      int pq = o.p * o.q;
      int img = (int)(pix / pq);
      int rem = (int)(pix - (long long)img * pq);
      int pp = rem / o.q, qq = rem - pp * o.q;
      int tap = (int)(kk / o.c), ch = (int)(kk - (long long)tap * o.c);
      int r = tap / o.ksize, s = tap - r * o.ksize;
      if (o.flip) { r = o.ksize - 1 - r; s = o.ksize - 1 - s; }
      int ih = pp * o.stride - o.pad + r, iw = qq * o.stride - o.pad + s;
      if (ih < 0 || ih >= o.h || iw < 0 || iw >= o.w) return 0.f;
      return psg_ld(o.ptr + (((long long)img * o.h + ih) * o.w + iw) * o.ld + ch);
    }
    case PSG_OP_DGRAD: {
      // rows: conv-input pixels (n, p, q); gathered tensor: dY NHWC [n, h, w, c] = conv output grid
      int pq = o.p * o.q;
      int img = (int)(i / pq);
      int rem = (int)(i - (long long)img * pq);
      int ih = rem / o.q, iw = rem - ih * o.q;
      int tap = (int)(k / o.c), ch = (int)(k - (long long)tap * o.c);
      int r = tap / o.ksize, s = tap - r * o.ksize;
      int oh = ih + o.pad - r, ow = iw + o.pad - s;
      if (oh < 0 || ow < 0 || (oh % o.stride) || (ow % o.stride)) return 0.f;
      oh /= o.stride; ow /= o.stride;
      if (oh >= o.h || ow >= o.w) return 0.f;
      return psg_ld(o.ptr + (((long long)img * o.h + oh) * o.w + ow) * o.ld + ch);
    }
  }
  return 0.f;
}

template <typename T>
__global__ void __launch_bounds__(kThreads) simt_gemm_kernel(OperandView<T> A, OperandView<T> B, long long M, long long N,
                                                            long long K, PsgEpilogue epi, long long k_per_split,
                                                            long long split_stride) {
  __shared__ float sA[TK][TM + 4];
  __shared__ float sB[TK][TN + 4];
  const int tid = threadIdx.x;
  const long long m0 = (long long)blockIdx.y * TM, n0 = (long long)blockIdx.x * TN;
  const int tx = tid & 15, ty = tid >> 4;  // 16 x 16 threads, 4x4 outputs each
  float acc[4][4];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;

  // loader mapping: K-contiguous operands want consecutive threads along k; MN-contiguous along rows
  const bool a_kfast = (A.mode == PSG_OP_KMAJOR || A.mode == PSG_OP_IM2COL || A.mode == PSG_OP_DGRAD);
  const bool b_kfast = (B.mode == PSG_OP_KMAJOR || B.mode == PSG_OP_IM2COL || B.mode == PSG_OP_DGRAD);

  // split-K: blockIdx.z owns k in [kbeg, kend) and writes its own fp32 partial [M, ldc] at out + z * split_stride
  const long long kbeg = (long long)blockIdx.z * k_per_split;
  const long long kend = (kbeg + k_per_split < K) ? kbeg + k_per_split : K;
  if (gridDim.z > 1) {
    A.K = kend;
    B.K = kend;
    epi.out = reinterpret_cast<float*>(epi.out) + (long long)blockIdx.z * split_stride;
  }
  for (long long k0 = kbeg; k0 < kend; k0 += TK) {
#pragma unroll
    for (int e = 0; e < (TM * TK) / kThreads; ++e) {
      int idx = tid + e * kThreads;
      int r, kk;
      if (a_kfast) { kk = idx % TK; r = idx / TK; } else { r = idx % TM; kk = idx / TM; }
      sA[kk][r] = fetch(A, m0 + r, k0 + kk);
    }
#pragma unroll
    for (int e = 0; e < (TN * TK) / kThreads; ++e) {
      int idx = tid + e * kThreads;
      int r, kk;
      if (b_kfast) { kk = idx % TK; r = idx / TK; } else { r = idx % TN; kk = idx / TN; }
      sB[kk][r] = fetch(B, n0 + r, k0 + kk);
    }
    __syncthreads();
#pragma unroll
    for (int kk = 0; kk < TK; ++kk) {
      float a[4], b[4];
#pragma unroll
      for (int i = 0; i < 4; ++i) a[i] = sA[kk][ty * 4 + i];
#pragma unroll
      for (int j = 0; j < 4; ++j) b[j] = sB[kk][tx * 4 + j];
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(a[i], b[j], acc[i][j]);
    }
    __syncthreads();
  }
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    long long m = m0 + ty * 4 + i;
    if (m >= M) continue;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      long long n = n0 + tx * 4 + j;
      if (n < N) psg_epilogue_scalar(epi, acc[i][j], m, n, N);
    }
  }
}

template <typename T>
static OperandView<T> make_view(const PsgOperand& o, long long rows, long long K) {
  OperandView<T> v;
  v.ptr = reinterpret_cast<const T*>(o.ptr);
  v.mode = o.mode; v.ld = o.ld; v.rows = rows; v.K = K;
  v.n = o.n; v.h = o.h; v.w = o.w; v.c = o.c; v.p = o.p; v.q = o.q;
  v.stride = o.stride > 0 ? o.stride : 1; v.pad = o.pad; v.ksize = o.ksize > 0 ? o.ksize : 1; v.flip = o.flip;
  return v;
}

}  // namespace simt

extern "C" {

int psg_simt_gemm(const PsgGemmDesc* d, void* stream) {
  using namespace simt;
  PSG_CHECK_ARG(d != nullptr, "psg_simt_gemm: null desc");
  PSG_CHECK_ARG(d->M > 0 && d->N > 0 && d->K >= 0, "psg_simt_gemm: bad sizes M=%lld N=%lld K=%lld", d->M, d->N, d->K);
  PSG_CHECK_ARG(d->a.ptr && d->b.ptr && d->epi.out, "psg_simt_gemm: null pointer");
  long long gy = (d->M + TM - 1) / TM, gx = (d->N + TN - 1) / TN;
  PSG_CHECK_ARG(gy <= 65535, "psg_simt_gemm: M too large for grid.y");
  // split-K (split_k > 1): fp32 partials [split][M][ldc] at epi.out, to be folded by psg_sum_partials
  int split = d->split_k > 1 ? d->split_k : 1;
  long long k_per_split = d->K;
  if (split > 1) {
    PSG_CHECK_ARG(d->epi.out_dtype == PSG_DTYPE_F32 && !d->epi.accumulate && !d->epi.residual, "psg_simt_gemm: split-K needs a plain fp32 epilogue");
    k_per_split = ((d->K + split - 1) / split + TK - 1) / TK * TK;
    split = (int)((d->K + k_per_split - 1) / k_per_split);
    PSG_CHECK_ARG(split == d->split_k, "psg_simt_gemm: split_k=%d does not divide K=%lld into TK-aligned slices", d->split_k, d->K);
  }
  const long long split_stride = d->M * d->epi.ldc;
  dim3 grid((unsigned)gx, (unsigned)gy, (unsigned)split);
  cudaStream_t s = (cudaStream_t)stream;
  if (d->in_dtype == PSG_DTYPE_F32) {
    simt_gemm_kernel<float><<<grid, kThreads, 0, s>>>(make_view<float>(d->a, d->M, d->K), make_view<float>(d->b, d->N, d->K),
                                                     d->M, d->N, d->K, d->epi, k_per_split, split_stride);
  } else if (d->in_dtype == PSG_DTYPE_BF16) {
    simt_gemm_kernel<__nv_bfloat16><<<grid, kThreads, 0, s>>>(make_view<__nv_bfloat16>(d->a, d->M, d->K),
                                                             make_view<__nv_bfloat16>(d->b, d->N, d->K), d->M, d->N, d->K,
                                                             d->epi, k_per_split, split_stride);
  } else {
    psg_set_error("psg_simt_gemm: bad in_dtype %d", d->in_dtype);
    return PSG_ERR_INVALID;
  }
  PSG_CHECK_LAUNCH("psg_simt_gemm");
  return PSG_OK;
}

}  // extern "C"
