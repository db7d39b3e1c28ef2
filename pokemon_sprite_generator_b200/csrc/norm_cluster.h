// Internal interface of norm_cluster.cu (cluster-split single-pass GroupNorm); called by the psg_groupnorm_fused_* entry
// points in norm_fused.cu.  Return PSG_OK, PSG_ERR_UNSUPPORTED (shape outside the plan: the caller falls back to the
// slab kernels) or PSG_ERR_CUDA.
#pragma once
#include <cuda_runtime.h>

// tunables: which = 0 fwd threads, 1 bwd threads, 2 fwd bytes of x per CTA, 3 largest cluster, 4 vectors per unit row
// (10 / 20 / 0 = by shape), 5 bwd bytes of x per CTA (0 = by shape), 6 L2 prefetch one residency ahead (0 off, 1 auto,
// > 1 resident CTAs assumed; env PSG_GN_PREFETCH sets the start value); value < 0 only reads
int gnc_tune(int which, int value);
int gnc_supported(int B, int HW, int C, int G, int bwd);
// out = {CC, S (cluster size), rows per CTA, R, TU, U, iters, smem bytes}
int gnc_plan(int B, int HW, int C, int G, int bwd, int* out);
int gnc_fwd(const void* x, long long ld_x, void* y, long long ld_y, const float* gamma, const float* beta, float* stats, int B,
            int HW, int C, int G, float eps, int act, cudaStream_t stream);
// main kernel only: writes partial[b][c] = {s1, s2, sum_pix dx}; the caller folds it over samples
int gnc_bwd(const void* dy, long long ld_dy, const void* x, long long ld_x, void* dx, long long ld_dx, const float* gamma,
            const float* beta, const float* stats, float* partial, float* dx_colsum, long long ld_colsum, int B, int HW, int C, int G,
            int act, int accumulate_dx, cudaStream_t stream);
