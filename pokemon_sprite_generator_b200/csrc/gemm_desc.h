// Plain-C description of one GEMM / implicit-GEMM convolution problem, shared by the two engines
// (tcgen05 `psg_umma_gemm`, SIMT `psg_simt_gemm`) and mirrored field-for-field by ctypes in
// pokemon_sprite_generator_b200/_lib.py.  All sizes in elements.
//
//   C[m, n] = epilogue( sum_k A(m, k) * B(n, k) )
//
// Operand access modes (how the logical [rows x K] operand is laid out in memory):
//   PSG_OP_KMAJOR   elem(i,k) = ptr[i*ld + k]                       (activations [tokens, C], weights [N, K])
//   PSG_OP_MNMAJOR  elem(i,k) = ptr[k*ld + i]                       (transposed view, used by wgrad / dgrad)
//   PSG_OP_IM2COL   i = output pixel (n,p,q), k = tap*C + c         (conv fprop / stride-1 dgrad, NHWC input)
//   PSG_OP_IM2COL_T i = tap*C + c,            k = output pixel      (conv wgrad "B" operand)
//   PSG_OP_DGRAD    i = input pixel (n,h,w),  k = tap*C + c over dY (general-stride dgrad gather; SIMT only)
//   PSG_OP_CONVW_T  i = cin,                  k = tap*Cout + cout     (conv weight [Cout][tap][Cin] read transposed: the
//                                                                      dgrad "B" operand; n = Cout, c = Cin; tcgen05 only)
#pragma once
#include "gemm_epilogue.cuh"

#define PSG_OP_KMAJOR 0
#define PSG_OP_MNMAJOR 1
#define PSG_OP_IM2COL 2
#define PSG_OP_IM2COL_T 3
#define PSG_OP_DGRAD 4
#define PSG_OP_CONVW_T 5

struct PsgOperand {
  const void* ptr;
  int mode;
  long long ld;      // row pitch (KMAJOR/MNMAJOR) or pixel pitch (conv modes), elements
  // conv geometry (conv modes only): the tensor being gathered is NHWC [n, h, w, c] with pixel pitch ld
  int n, h, w, c;
  int p, q;          // spatial size of the "row" index space (IM2COL*: conv output; DGRAD: conv input)
  int stride, pad, ksize;
  int flip;          // IM2COL only: bit 0: use tap (ksize-1-r, ksize-1-s) -> stride-1 dgrad; bits 8..15: 0, or (pad_hi + 1) when the
                     // bottom / right padding differs from `pad` (tcgen05 engine only)
};

struct PsgGemmDesc {
  PsgOperand a, b;
  long long M, N, K;
  int in_dtype;      // dtype of A and B (PSG_DTYPE_*)
  int split_k;       // >1: write fp32 partials [split][M][ldc] to epi.out (epilogue must be plain fp32)
  PsgEpilogue epi;
};
