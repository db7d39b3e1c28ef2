// Internal interface of norm_stream.cu (streaming two-phase GroupNorm backward for tensors larger than L2); called by
// psg_groupnorm_fused_bwd_ws in norm_fused.cu.  Return PSG_OK, PSG_ERR_UNSUPPORTED (shape outside the plan or workspace too
// small: the caller falls back to the cluster / slab kernels) or PSG_ERR_CUDA.
#pragma once
#include <cuda_runtime.h>

// tunables: which = 0 bytes of x + dy per L2-resident sample group, 1 target pixel rows per chunk (0 = by shape), 2 smallest
// tensor (bytes of x) the auto mode routes here; value < 0 only reads
long long gns_tune(int which, long long value);
long long gns_workspace_floats(int B, int HW, int C, int G);       // 0 = unsupported shape
int gns_wants(int B, int HW, int C);                               // auto mode: tensor at least as large as tunable 2
// out = {vectors per row, row lanes, threads, rows per chunk, chunks per sample, samples per group, smem bytes, grid}
int gns_plan(int B, int HW, int C, int G, int* out);
// writes workspace[0 .. B*C*3) = partial[b][c] = {s1, s2, sum_pix dx}; the caller folds it over samples
int gns_bwd(const void* dy, long long ld_dy, const void* x, long long ld_x, void* dx, long long ld_dx, const float* gamma,
            const float* beta, const float* stats, float* workspace, long long workspace_floats, float* dx_colsum, long long ld_colsum,
            int B, int HW, int C, int G, int act, int accumulate_dx, cudaStream_t stream);
int gns_timeout_flag();
