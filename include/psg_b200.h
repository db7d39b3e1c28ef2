/* psg_b200 -- C ABI of the B200-native latent-diffusion hot path (libpsg_b200.so, sm_100a only).
 *
 * The reference (GabrieleConte/pokemon-sprite-generator) is pure Python: it has no FFI of its own.  Its "operator API"
 * for this path is the nn.Module / trainer contract; the Python host side
 * (pokemon_sprite_generator_b200/{unet,scheduler,trainer}.py) mirrors that contract and binds THIS header through
 * ctypes.  Each entry point below names the reference op(s) it replaces (paths relative to the reference root).
 *
 * Conventions: every function returns 0 (PSG_OK) or a negative error code and never allocates or synchronises (the test
 * hooks psg_*_timeout_flag excepted); `stream` is a cudaStream_t.  Host-side state: a process-wide kernel-launch counter; per
 * DEVICE (the calling thread's current device) the registered stream-K workspace and the SM reservation of the tcgen05 GEMM
 * engine -- one stream at a time may run stream-K GEMMs on a device; and the process-wide measurement / test hooks named as
 * such below (psg_umma_pairs, psg_umma_debug, psg_attn_fused_split, psg_attn_umma_enable, psg_groupnorm_* tuning); all pointers are device pointers
 * unless noted; the caller owns all memory (workspaces included); psg_last_error() describes the last failure of the
 * calling thread.  dtype: 0 = fp32, 1 = bf16 (activation storage type).  Token-major tensors are [rows, C] with row
 * pitch `ld` in elements (channel slices of wider buffers are valid operands).
 */
#ifndef PSG_B200_H
#define PSG_B200_H

#include <stddef.h>

#ifdef __cplusplus
extern "C" {
#endif

#define PSG_OK 0
#define PSG_ERR_INVALID -1
#define PSG_ERR_CUDA -2
#define PSG_ERR_UNSUPPORTED -3

#define PSG_DTYPE_F32 0
#define PSG_DTYPE_BF16 1
#define PSG_ACT_NONE 0
#define PSG_ACT_GELU 1
#define PSG_ACT_SILU 2
#define PSG_ACT_MUL 3   /* aux_in only: multiply by the stored value (a derivative saved by the forward epilogue) */
#define PSG_ACT_RELU 4  /* forward only (VAE encoder stem) */
#define PSG_OP_KMAJOR 0
#define PSG_OP_MNMAJOR 1
#define PSG_OP_IM2COL 2
#define PSG_OP_IM2COL_T 3
#define PSG_OP_DGRAD 4
#define PSG_OP_CONVW_T 5   /* conv weight [Cout][tap][Cin] read as B(n = cin, k = tap*Cout + cout): dgrad without a transposed copy */

/* ---- library ------------------------------------------------------------------------------------------------- */
int psg_version(void);
const char* psg_last_error(void);
int psg_check_device(void);                 /* 0 iff the current device is compute capability 10.x */
long long psg_launch_count(int reset);      /* kernels launched by this library so far */
int psg_umma_timeout_flag(void);            /* test hook: 1 if a tcgen05 pipeline wait timed out (synchronises) */

/* ---- fused epilogue + GEMM / implicit-GEMM convolution (csrc/gemm_epilogue.cuh, csrc/gemm_desc.h) --------------
 * Replaces aten::conv2d / aten::addmm forward and backward behind nn.Conv2d, nn.Linear and the MHA projections:
 *   src/models/unet.py:29-33 (time MLP), :80,90,96 (ResBlock convs + 1x1 skip), :83,86 (time/text proj),
 *   :160-187 (attention in/out projections, text_proj, FFN), :325-399 (init/down/up/final convs).
 *   v = acc + bias[n] + rowbias[(m / rows_per_group) * ld_rowbias + n]; aux_out = v; v = act(v);
 *   v *= act'(aux_in); v = dropout(v); out = alpha * v + residual (+ out if accumulate)                              */
typedef struct PsgEpilogue {
  void* out; long long ldc; int out_dtype; int act_dtype;
  const float* bias; const float* rowbias; int rows_per_group; long long ld_rowbias;
  int act; float alpha;
  const void* residual; long long ldr;
  void* aux_out; const void* aux_in; long long ld_aux; int aux_act;
  int accumulate;
  unsigned long long drop_seed; unsigned int drop_threshold; float drop_scale;
} PsgEpilogue;

typedef struct PsgOperand {
  const void* ptr; int mode; long long ld;
  int n, h, w, c;          /* NHWC geometry of the gathered tensor (conv modes) */
  int p, q;                /* spatial size of the row index space */
  int stride, pad, ksize, flip;
} PsgOperand;

typedef struct PsgGemmDesc {
  PsgOperand a, b;         /* C[m,n] = epilogue(sum_k A(m,k) * B(n,k)) */
  long long M, N, K;
  int in_dtype;
  int split_k;             /* SIMT engine only: >1 = fp32 partial planes [split][M][ldc] (the tcgen05 engine is stream-K) */
  PsgEpilogue epi;
} PsgGemmDesc;

/* tcgen05 / TMEM / TMA engine (bf16 operands, fp32 accumulate), persistent CTAs.  block_n: 0 = auto, or 64/128/160/256;
 * m_tiles: 0 = auto, 1 or 2 (CTA tile = 128*m_tiles rows).  psg_umma_plan reports the automatic choice. */
int psg_umma_gemm(const PsgGemmDesc* desc, int block_n, void* stream);
int psg_umma_gemm_ex(const PsgGemmDesc* desc, int block_n, int m_tiles, void* stream);
/* As psg_umma_gemm_ex, naming the stream-K workspace lane (0 or 1): launches that may be in flight together on different streams
 * of one device must use different lanes (psg_umma_gemm / _ex use lane 0). */
int psg_umma_gemm_lane(const PsgGemmDesc* desc, int block_n, int m_tiles, int lane, void* stream);
int psg_umma_plan(const PsgGemmDesc* desc, int* block_n, int* m_tiles);
/* Stream-K scheduling: tiles whose k-range is shared between CTAs exchange fp32 partial accumulators through a
 * caller-owned workspace (psg_umma_workspace_bytes() bytes, 256B aligned, first 1 KiB zeroed), registered once. */
size_t psg_umma_workspace_bytes(void);
int psg_umma_set_workspace(void* workspace, size_t bytes);
int psg_umma_pairs(int on);                    /* cta_group::2 CTA pairs: 0 never, 1 where measured to pay (default), 2 always */
int psg_umma_max_pairs(void);                  /* co-resident 2-CTA clusters on this device */
int psg_umma_reserve_sms(int n);               /* SMs the persistent GEMM grids leave to a concurrent collective (n < 0: read); returns the previous value */
int psg_umma_debug(int flags);                 /* profiling aid: 1 = skip the epilogue body (mainloop time alone); 0 = normal */
/* CUDA-core fp32-accumulate engine (fp32 parity mode, edge shapes, general-stride dgrad gather). */
int psg_simt_gemm(const PsgGemmDesc* desc, void* stream);

/* ---- diffusion element-wise kernels ------------------------------------------------------------------------------
 * psg_q_sample          NoiseScheduler.add_noise + clamp  src/training/improved_diffusion_trainer.py:50-65,363
 * psg_smooth_l1_fwd_bwd nn.SmoothL1Loss(beta) + backward  src/training/improved_diffusion_trainer.py:300,388,396
 * psg_ddpm_step         mode 0: ddpm_sample update        src/training/improved_diffusion_trainer.py:543-567
 *                       mode 1: sample_previous_timestep  src/training/final_trainer.py:52-71                        */
int psg_q_sample(const float* x0, const float* noise, const long long* t, const float* sqrt_ac, const float* sqrt_1mac,
                 float* out, int batch, int n_per, int num_t, int do_clamp, float clamp_lo, float clamp_hi, int* nonfinite,
                 void* stream);
int psg_smooth_l1_fwd_bwd(const float* pred, const float* target, float* dpred, float* loss, void* workspace, long long n,
                          float beta, float grad_scale, void* stream);
int psg_ddpm_step(const float* x, const float* eps, const float* z, float* out, long long n, int mode, const float* tab0,
                  const float* tab1, const float* tab2, const float* tab3, int t, int num_t, void* stream);

/* VAE reparameterisation out = mu + eps * exp(0.5 * logvar) [clamped to [lo, hi]]: src/models/vae_decoder.py:119-123 followed by the
 * trainer's clamp (src/training/improved_diffusion_trainer.py:363); the latent cache samples with it every epoch */
int psg_reparam(const float* mu, const float* logvar, const float* eps, float* out, long long n, int do_clamp, float lo, float hi,
                void* stream);

/* the reference's two other reverse-step formulas: mode 2 = src/training/diffusers_trainer.py:76-100, mode 3 = gradio_app.py:324-361;
 * coef5 is a HOST array of the per-step scalars evaluated as the reference evaluates them (see csrc/diffusion_ops.cu) */
int psg_reverse_step(const float* x, const float* eps, const float* z, float* out, long long n, int mode, const float* coef5,
                     void* stream);

/* ---- fused tensor-core attention (bf16): scores stay on chip, forward saves only the row log-sum-exp ------------------
 * replaces the nn.MultiheadAttention core, src/models/unet.py:160-173,217,235 (softmax over keys, dropout on probabilities) */
int psg_attn_fused_ok(int B, int H, int Lq, int Lk, int hd);
int psg_attn_fused_small_bwd(int on); /* test / measurement hook: single-kernel backward for Lq, Lk <= 64 (default 1); returns the previous value */
int psg_attn_fused_split(int n);   /* test / measurement hook: CTAs per (batch, head), 0 = by problem size; returns the previous value */
int psg_attn_fused_fwd(const void* q, long long ldq, const void* k, long long ldk, const void* v, long long ldv, void* o, long long ldo,
                       float* lse, int B, int H, int Lq, int Lk, int hd, float scale, unsigned long long drop_seed, float drop_p,
                       void* stream);
int psg_attn_fused_bwd(const void* q, long long ldq, const void* k, long long ldk, const void* v, long long ldv, const void* o,
                       long long ldo, const void* dout, long long lddo, const float* lse, float* delta, void* dq, long long lddq,
                       void* dk, long long lddk, void* dv, long long lddv, int B, int H, int Lq, int Lk, int hd, float scale,
                       unsigned long long drop_seed, float drop_p, void* stream);

/* tcgen05 / TMEM attention (csrc/attention_umma.cu): every product on the 5th-generation tensor cores, scores in TMEM.  Takes the
 * problems with 65..256 queries, <= 256 keys, head_dim % 16 == 0 and <= 256; psg_attn_fused_* route to it when psg_attn_umma_ok. */
int psg_attn_umma_enable(int on);  /* test / measurement hook: 0 = never route to these kernels, 1 = default; returns the previous value */
int psg_attn_umma_ok(int B, int H, int Lq, int Lk, int hd);
int psg_attn_umma_timeout_flag(void); /* test hook: 1 if a bounded barrier wait of these kernels expired since the last call */
int psg_attn_umma_fwd(const void* q, long long ldq, const void* k, long long ldk, const void* v, long long ldv, void* o, long long ldo,
                      float* lse, int B, int H, int Lq, int Lk, int hd, float scale, unsigned long long drop_seed, float drop_p,
                      void* stream);
int psg_attn_umma_bwd(const void* q, long long ldq, const void* k, long long ldk, const void* v, long long ldv, const void* o,
                      long long ldo, const void* dout, long long lddo, const float* lse, float* delta, void* dq, long long lddq,
                      void* dk, long long lddk, void* dv, long long lddv, int B, int H, int Lq, int Lk, int hd, float scale,
                      unsigned long long drop_seed, float drop_p, void* stream);

/* ---- text encoder pieces (src/models/text_encoder.py:137-163: BertModel -> projection -> LayerNorm) ----------------------------- */
int psg_layernorm(const void* x, long long ldx, void* y, long long ldy, const float* gamma, const float* beta, long long rows, int D,
                  float eps, int in_dtype, int out_dtype, void* stream);
int psg_bert_embed(const long long* ids, const long long* type_ids, const float* word, const float* pos, const float* type, float* out,
                   long long rows, int L, int D, int vocab, void* stream);
/* psg_attn_fwd with a per-sample number of valid keys (BERT's padding mask); key_len: device int32 [B], nullable */
int psg_attn_fwd_keylen(const void* q, long long ldq, const void* k, long long ldk, const void* v, long long ldv, void* o,
                        long long ldo, float* lse, const int* key_len, int B, int H, int Lq, int Lk, int hd, float scale, int dtype,
                        unsigned long long drop_seed, float drop_p, void* stream);

/* ---- GroupNorm (+SiLU)  nn.GroupNorm + F.silu, src/models/unet.py:79,89,115,127,156-157,214,231,397-398 ---------- */
int psg_groupnorm_slices(int B, int HW);
int psg_groupnorm_fwd(const void* x, long long ld_x, void* y, long long ld_y, const float* gamma, const float* beta,
                      float* stats, float* workspace, int B, int HW, int C, int G, float eps, int act, int dtype, void* stream);
int psg_groupnorm_bwd(const void* dy, long long ld_dy, const void* x, long long ld_x, void* dx, long long ld_dx,
                      const float* gamma, const float* beta, const float* stats, float* dgamma, float* dbeta, float* workspace,
                      int B, int HW, int C, int G, int act, int dtype, int accumulate_dx, int accumulate_params, void* stream);

/* single-pass variants (bf16; the (sample, channel-chunk) slab is staged once in shared memory: 2N / 3N bytes of traffic).
 * The backward optionally emits sum-over-pixels of dx per (sample, channel) [B, ld_colsum] and per channel [C]: the
 * gradients of the conv bias and of the broadcast time/text conditioning that produced x (unet.py:116-124).           */
int psg_groupnorm_fused_ok(int B, int HW, int C, int G, int dtype);
int psg_groupnorm_fused_plan(int B, int HW, int C, int G, int* out8);
/* The fused entry points run the cluster-split kernels (norm_cluster.cu: pixels of a (sample, 160-channel) unit split over
 * the CTAs of a thread-block cluster, partial sums exchanged through distributed shared memory) where their plan applies
 * and the slab kernels otherwise.  psg_groupnorm_fused_mode(1) forces the slab kernels (A/B measurements); returns the
 * previous mode (modes 2 / 3: see psg_groupnorm_fused_bwd_ws).  psg_groupnorm_cluster_plan: out8 = {CC, cluster size, rows per CTA, R, TU, U, iters, smem bytes}.     */
int psg_groupnorm_fused_mode(int mode);
int psg_groupnorm_cluster_plan(int B, int HW, int C, int G, int bwd, int* out8);
int psg_groupnorm_cluster_tune(int which, int value); /* measurement hook: 0 fwd threads, 1 bwd threads, 2 fwd slab bytes per CTA, 3 max cluster, 4 vectors per unit row, 5 bwd slab bytes, 6 L2 prefetch one residency ahead (0 off / 1 auto / resident CTAs) */
int psg_groupnorm_fused_fwd(const void* x, long long ld_x, void* y, long long ld_y, const float* gamma, const float* beta,
                            float* stats, int B, int HW, int C, int G, float eps, int act, void* stream);
int psg_groupnorm_fused_bwd(const void* dy, long long ld_dy, const void* x, long long ld_x, void* dx, long long ld_dx,
                            const float* gamma, const float* beta, const float* stats, float* dgamma, float* dbeta,
                            float* workspace, float* dx_colsum, long long ld_colsum, float* bias_total, int B, int HW, int C, int G,
                            int act, int accumulate_dx, int accumulate_params, void* stream);
/* The same with the workspace size stated: >= psg_groupnorm_bwd_workspace_floats(...) floats lets the backward of tensors beyond
 * the L2 (27x27 / 14x14 levels at batch 256) run as the streaming two-phase kernel (norm_stream.cu: per-chunk raw moments,
 * then dx out of an L2-resident sample group, one launch); psg_groupnorm_fused_bwd == this with B*C*3 floats.
 * psg_groupnorm_fused_mode: 0 auto, 1 slab kernels only, 2 streaming backward whenever its plan applies, 3 auto without it.
 * psg_groupnorm_stream_tune: 0 bytes of x + dy per L2 group, 1 pixel rows per chunk, 2 smallest tensor (bytes of x) routed
 * to it.  psg_groupnorm_timeout_flag: test hook, 1 if a bounded in-kernel wait expired since the last call (synchronises). */
int psg_groupnorm_fused_bwd_ws(const void* dy, long long ld_dy, const void* x, long long ld_x, void* dx, long long ld_dx,
                               const float* gamma, const float* beta, const float* stats, float* dgamma, float* dbeta,
                               float* workspace, long long workspace_floats, float* dx_colsum, long long ld_colsum, float* bias_total,
                               int B, int HW, int C, int G, int act, int accumulate_dx, int accumulate_params, void* stream);
long long psg_groupnorm_bwd_workspace_floats(int B, int HW, int C, int G);
long long psg_groupnorm_stream_tune(int which, long long value);
int psg_groupnorm_stream_plan(int B, int HW, int C, int G, int* out8);
int psg_groupnorm_timeout_flag(void);

/* ---- attention core  nn.MultiheadAttention(batch_first) softmax(QK^T/sqrt(d))V, src/models/unet.py:160-173,217,235 */
int psg_attn_fwd(const void* q, long long ldq, const void* k, long long ldk, const void* v, long long ldv, void* o,
                 long long ldo, float* lse, int B, int H, int Lq, int Lk, int hd, float scale, int dtype,
                 unsigned long long drop_seed, float drop_p, void* stream);
int psg_attn_bwd(const void* q, long long ldq, const void* k, long long ldk, const void* v, long long ldv, const void* o,
                 long long ldo, const void* dout, long long lddo, const float* lse, float* dsum, void* dq, long long lddq,
                 void* dk, long long lddk, void* dv, long long lddv, int B, int H, int Lq, int Lk, int hd, float scale,
                 int dtype, unsigned long long drop_seed, float drop_p, void* stream);

/* tensor-core attention pieces (bf16 path): batched C[b,h] = alpha * A[b,h] B[b,h]^T with either operand stored
 * transposed (mma.sync m16n8k16), and the row softmax / its backward with the dropout mask of the probabilities. */
int psg_bmm_bf16(const void* a, long long a_sb, long long a_sh, long long lda, int trans_a, const void* b, long long b_sb,
                 long long b_sh, long long ldb, int trans_b, void* c, long long c_sb, long long c_sh, long long ldc, int c_is_f32,
                 int batch, int heads, int M, int N, int K, float alpha, void* stream);
int psg_softmax_fwd(const float* S, void* P, void* Pd, long long rows, int Lk, int ldp, unsigned long long seed, float drop_p,
                    void* stream);
int psg_softmax_bwd(const void* P, const float* dPd, void* dS, void* Pd, long long rows, int Lk, int ldp, unsigned long long seed,
                    float drop_p, void* stream);

/* ---- layout, resize, reductions, conditioning inputs, weight packing ---------------------------------------------
 * upsample: nn.Upsample(bilinear, align_corners=False) unet.py:365,375,385; timestep embedding unet.py:47-50;
 * mean pool: AdaptiveAvgPool1d(1) unet.py:322,445; copy_strided: torch.cat unet.py:482-503; colsum: bias gradients. */
/* VAE decoder (src/models/vae_decoder.py:33-65,128-222): per-sample transpose (the raw-reshape K / V of its cross-attention), tanh */
int psg_batched_transpose(const void* in, void* out, int B, int R, int Cc, int dtype, void* stream);
int psg_tanh(const float* x, float* y, long long n, void* stream);
int psg_nchw_to_tokens(const float* src, void* dst, long long ld, int B, int C, int HW, int dtype, void* stream);
int psg_tokens_to_nchw(const void* src, long long ld, float* dst, int B, int C, int HW, int dtype, void* stream);
int psg_copy_strided(const void* src, long long lds, void* dst, long long ldd, long long rows, int C, int accumulate,
                     int dtype, void* stream);
int psg_colsum_slices(int groups, int rows_per_group);
int psg_colsum(const void* x, long long ld, int groups, int rows_per_group, int C, float* out_groups, long long ld_groups,
               int acc_groups, float* out_total, int acc_total, float scale, float* workspace, int dtype, void* stream);
int psg_upsample_bilinear_fwd(const void* x, long long ldx, void* y, long long ldy, int B, int C, int IH, int IW, int OH,
                              int OW, int dtype, void* stream);
int psg_upsample_bilinear_bwd(const void* dy, long long lddy, void* dx, long long lddx, int B, int C, int IH, int IW, int OH,
                              int OW, int accumulate, int dtype, void* stream);
/* stride-2 3x3 dgrad by output parity: four 2x2 / 1x1 stride-1 convolutions over dY (class weights from the dgrad layout
 * [Cin][9][Cout]), then one interleave of the four [B, P, Q, C] class results into dX [B, H, W, C] */
int psg_dgrad_s2_weights(const void* wd, void* w00, void* w01, void* w10, void* w11, int Cin, int Cout, int dtype, void* stream);
int psg_interleave2x2(const void* c00, const void* c01, const void* c10, const void* c11, void* out, long long ldo, int B, int C, int P,
                      int Q, int H, int W, int accumulate, int dtype, void* stream);
int psg_dilate2(const void* dy, long long lddy, void* out, long long ldo, int B, int C, int P, int Q, int H, int W, int dtype,
                void* stream);
int psg_dropout_scale(const void* x, long long ldx, void* out, long long ldo, long long rows, int C, float alpha,
                      unsigned long long seed, float drop_p, int dtype, void* stream);
int psg_timestep_embedding(const long long* t, const float* coeff, float* out, int B, int half, void* stream);
int psg_mean_pool(const float* x, float* out, int B, int L, int D, void* stream);
int psg_pack_conv_weight(const float* w_oihw, void* wp, void* wd, int Cout, int Cin, int kk, int Cin_p, int Cout_p, int dtype,
                         void* stream);
/* every tap-major bf16 conv weight [Cout][kk][Cin] -> dgrad layout [Cin][kk][Cout] in one launch; jobs = n x {src offset, dst offset,
 * Cout, Cin, kk} (elements), host array */
int psg_conv_weights_transpose(const void* src_base, void* dst_base, const long long* jobs, int n, void* stream);
int psg_pack_linear_weight(const float* w, void* wk, void* wt, int N, int K, int dtype, void* stream);
int psg_wgrad_finalize(const float* partial, int splits, long long split_stride, float* grad_oihw, int Cout, int Cin, int kk,
                       int Cin_p, int accumulate, void* stream);
int psg_sum_partials(const float* partial, int splits, long long split_stride, float* out, long long n, int accumulate,
                     void* stream);

/* ---- optimiser  (478 x grad.norm().item() + clip_grad_norm_ + AdamW(eps=1e-6).step(),
 *                  src/training/improved_diffusion_trainer.py:277-283,399-413) ------------------------------------- */
int psg_sumsq(const float* x, long long n, float* out_sumsq, int accumulate, void* workspace, void* stream);
int psg_clip_coef(const float* sumsq, float max_norm, float* state /* [3]: norm, coef, finite */, void* stream);
/* betas are doubles, as torch holds them: the kernel uses float(beta) and float(1 - beta), the latter formed in double */
int psg_adamw_step(float* p, const float* g, float* m, float* v, long long n, float lr, double beta1, double beta2, float eps,
                   float weight_decay, long long step, const float* state, void* bf16_shadow /* nullable: bf16 copy of p */,
                   void* stream);
/* generalisation: step <= 0 takes the applied-step count from state[3] (written by psg_clip_coef_count, which skips
 * non-finite steps like the reference's `continue`, :395-397); coupled_l2 = 1 is torch.optim.Adam(weight_decay) (:285-292) */
int psg_clip_coef_count(const float* sumsq, float max_norm, float* state4 /* [4]: norm, coef, finite, applied steps */,
                        void* stream);
int psg_adam_step(float* p, const float* g, float* m, float* v, long long n, float lr, double beta1, double beta2, float eps,
                  float weight_decay, long long step, int coupled_l2, const float* state, void* bf16_shadow, void* stream);
int psg_cast_bf16(const float* x, void* y_bf16, long long n, void* stream);
int psg_scale_inplace(float* x, long long n, const float* state, float extra, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* PSG_B200_H */
