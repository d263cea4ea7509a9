/* fedcola_b200 — C ABI of the B200-native FedCola round hot path (libfedcola_b200.so).
 *
 * The reference (imguangyu/FedCola) is pure Python: its "plugin API" is name resolution of
 * {Algorithm}Server / {Algorithm}Client / mome_* model factories.  This header is the
 * boundary a maintainer binds underneath those classes (ctypes stub in INTEGRATION.md).  Every entry
 * point:
 *   - is `extern "C"`, takes plain device pointers + sizes, an explicit CUDA device ordinal and a
 *     `cudaStream_t` (passed as void*), and returns 0 or a negative FC_ERR_* code;
 *   - allocates nothing persistent: all buffers are caller-owned (torch-allocated) device memory;
 *   - is re-entrant (callers are the reference's ThreadPoolExecutor workers, fedavgserver.py:566).
 * `fc_last_error()` returns the calling thread's last error text.
 *
 * Unless stated otherwise, "ref:" cites the file:line under /root/reference that the call replaces.
 */
#ifndef FEDCOLA_B200_H_
#define FEDCOLA_B200_H_

#ifdef __cplusplus
extern "C" {
#endif

#define FC_ABI_VERSION 1

/* return codes of every int-returning entry point */
#define FC_OK 0
#define FC_ERR_INVALID (-1)      /* bad argument (validated before touching the device) */
#define FC_ERR_CUDA (-2)         /* a CUDA runtime / driver call failed; fc_last_error() has the text */
#define FC_ERR_UNSUPPORTED (-3)  /* combination not built / hardware path absent */

const char* fc_last_error(void);
int fc_abi_version(void);
/* number of CUDA kernels this library has launched so far in this process */
unsigned long long fc_launch_count(void);
/* Test / measurement hook: cap the CTA count of the persistent kernels (fc_gemm_bf16, fc_attention_fwd/bwd) so that
 * small problems run many tiles / items per CTA (the regime of the bench shapes).  0 = no cap.  Process-wide;
 * returns the previous value. */
int fc_set_grid_cap(int max_ctas);

/* Enable peer (NVLink) access from `device` to memory on `peer_device`, so that fc_aggregate launched on `device`
 * can read client arenas trained on another GPU of the same process in place (ref: the thread-per-client
 * `cuda:(i % ngpu)` placement, src/server/fedavgserver.py:310-311).  FC_ERR_UNSUPPORTED if there is no peer path. */
int fc_enable_peer_access(int device, int peer_device);

/* Host -> device gather of one batch: dst[i, :] = src_host[idx[i], :] (row_bytes each) as cudaMemcpyAsync on `stream`,
 * consecutive indices merged; src_host should be pinned, idx is a HOST array.
 * ref: src/client/fedavgclient.py:79-84 (DataLoader collation + .to(device) of every batch). */
int fc_h2d_rows(void* dst, const void* src_host, const long long* idx, int n, long long row_bytes, int device,
                void* stream);

/* ------------------------------------------------------------------------------------------------
 * Server aggregation                                    ref: src/server/fedavgserver.py:597,656-666
 *                                                            src/client/fedavgclient.py:158-184 (aux merge)
 * ------------------------------------------------------------------------------------------------ */
#define FC_AGG_MAX_OUT 4   /* outputs (global models) one parameter name can feed */
#define FC_AGG_LERP 0      /* sequential lerp, bit-exact with the reference (1 GPU, pinned order) */
#define FC_AGG_WSUM 1      /* closed-form weighted sum (multi-GPU partial sums) */
/* kinds of source entries */
#define FC_AGG_SRC_PLAIN 0 /* x = *src; fold x into the outputs with coef[]                          */
#define FC_AGG_SRC_HOLD 1  /* W of an aux-merged upload: keep, the next entry (MERGE) completes it   */
#define FC_AGG_SRC_MERGE 2 /* x = W_held + (*src) * (*scale_ptr); fold x with coef[]                 */

/* Floats covered by one tile of the plan (the host planner needs it to build job_tile_start). */
int fc_aggregate_tile_floats(void);

/* One launch aggregates every (global model, parameter) pair of the round.
 *   job j = one parameter name; covers tiles [job_tile_start[j], job_tile_start[j+1]) of
 *   fc_aggregate_tile_floats() floats each; job_numel[j] floats; job_nout[j] outputs whose old/new
 *   global segments are job_gin/job_gout[j*MAX_OUT + o] (device addresses; may alias);
 *   source entries [job_src_start[j], job_src_start[j+1]) of src_ptr/src_flag/scale_ptr/coef list the
 *   contributing clients in ascending client id; a client that uploads an aux-merged weight
 *   (W + A*s) contributes a HOLD entry (W) followed by a MERGE entry (A, s = *scale_ptr).
 *   coef[k*MAX_OUT+o] is c (LERP; 0 = skip) or the closed-form weight (WSUM);
 *   job_gscale[j*MAX_OUT+o] is the weight of the old global (WSUM only).
 * All table pointers are DEVICE pointers.  grid_ctas <= 0 selects 8 CTAs per SM. */
int fc_aggregate(int mode, int n_jobs, int n_tiles, const int* job_tile_start,
                 const long long* job_numel, const int* job_nout,
                 const unsigned long long* job_gin, const unsigned long long* job_gout,
                 const float* job_gscale, const int* job_src_start,
                 const unsigned long long* src_ptr, const int* src_flag,
                 const unsigned long long* scale_ptr, const float* coef, int grid_ctas,
                 int device, void* stream);

/* ------------------------------------------------------------------------------------------------
 * tcgen05 GEMM with fused epilogues            ref: src/models/mome.py:58-60 (W + s*A linear),
 *   :112-121 (fc1/GELU/fc2), :143-166 (qkv/proj), :252-265 (PatchEmbed conv), and their autograd.
 *   out[M,N] (+)= A[M,K] * B[N,K]^T, bf16 operands, fp32 accumulation in TMEM.
 *   a_mn_major / b_mn_major = 1: the operand is stored [K, rows] (rows contiguous) instead of [rows, K].
 * ------------------------------------------------------------------------------------------------ */
#define FC_EPI_BF16 0        /* out(bf16) = acc + bias                                              */
#define FC_EPI_GELU 1        /* x = acc + bias: out(bf16) = gelu_erf'(x) ; out2(bf16) = gelu_erf(x)  */
#define FC_EPI_RESID 2       /* out(f32)  = resid + row_scale[row/rows_per_group] * (acc + bias)    */
#define FC_EPI_MULAUX 3      /* out(bf16) = acc * aux (aux = bf16, e.g. the saved gelu'); colsum += column sums */
#define FC_EPI_F32 4         /* out(f32)  = acc + bias                                              */
#define FC_EPI_ATOMIC_F32 5  /* out(f32) += alpha * acc  (red.add; split-K; splits <= 0 = choose automatically) */
#define FC_EPI_PATCH 6       /* out(f32)[b*(P+1)+1+t] = acc + bias + pos[1+t]  (row = b*P + t)      */

int fc_gemm_bf16(int M, int N, int K, const void* A, long long lda, int a_mn_major, const void* B,
                 long long ldb, int b_mn_major, int epi, void* out, void* out2, long long ldo,
                 const float* bias, const float* resid, const float* row_scale, int rows_per_group,
                 const void* aux, const float* pos, int patches, float alpha, int splits, float* colsum,
                 int device, void* stream);

/* The same GEMM (shape, strides, majors, epilogue) for `groups` (<= FC_GEMM_MAX_GROUPS) independent operand sets in ONE
 * launch — the same layer of several clients that train in lockstep (ref: the ThreadPoolExecutor of
 * src/server/fedavgserver.py:566-577 runs these clients side by side, one kernel stream each).  Every pointer
 * argument of fc_gemm_bf16 becomes a host array of `groups` pointers (optional operands: a NULL table or NULL entries,
 * consistently across groups). */
#define FC_GEMM_MAX_GROUPS 4
int fc_gemm_bf16_grouped(int groups, int M, int N, int K, const void* const* A, long long lda, int a_mn_major,
                         const void* const* B, long long ldb, int b_mn_major, int epi, void* const* out,
                         void* const* out2, long long ldo, const float* const* bias, const float* const* resid,
                         const float* const* row_scale, int rows_per_group, const void* const* aux,
                         const float* const* pos, int patches, float alpha, int splits, float* const* colsum,
                         int device, void* stream);

/* fp32-accurate GEMM on the bf16 tensor pipe (validation mode, precision = 'fp32'): operands are (hi, lo) bf16 pairs,
 * X = X_hi + X_lo, X_lo = bf16(X - X_hi); acc = A_hi B_hi + A_hi B_lo + A_lo B_hi (relative error ~2^-16, better than
 * kind::tf32's 2^-10).  fp32-output epilogues only (F32, RESID, ATOMIC_F32, PATCH).
 * ref: the reference's strict-fp32 F.linear (src/models/mome.py:58-60,112-121,143-166). */
int fc_gemm_split(int M, int N, int K, const void* A_hi, const void* A_lo, long long lda, int a_mn_major,
                  const void* B_hi, const void* B_lo, long long ldb, int b_mn_major, int epi, void* out, long long ldo,
                  const float* bias, const float* resid, const float* row_scale, int rows_per_group, const float* pos,
                  int patches, float alpha, int splits, int device, void* stream);

/* Per-launch GEMM timing with CUDA events on the launching stream (measurement aid; off by default). */
void fc_gemm_profile(int enable);
long long fc_gemm_profile_collect(double* total_ms, double* total_flops);

/* ------------------------------------------------------------------------------------------------
 * Fused short-sequence attention (head_dim 64)        ref: src/models/mome.py:153-165 + autograd
 *   qkv  bf16 [B, N, 3, H, 64]  (output of the qkv Linear, untouched layout)
 *   out  bf16 [B, N, H*64]      lse fp32 [B, H, N] (row log-sum-exp of the scaled scores; may be NULL in fwd)
 * ------------------------------------------------------------------------------------------------ */
int fc_attention_fwd(const void* qkv, void* out, float* lse, int B, int N, int H, int head_dim, int device,
                     void* stream);
/* dbias (nullable): fp32 [3*H*64] += the qkv bias gradient: column sums of the stored dQ and dV thirds; the K third is
 * left untouched — it is identically zero because softmax is invariant to a shift of the scores */
int fc_attention_bwd(const void* qkv, const void* out, const void* d_out, const float* lse, void* dqkv,
                     float* dbias, int B, int N, int H, int head_dim, int device, void* stream);
/* The same attention for `groups` (<= FC_ATTN_MAX_GROUPS) clients' tensors of one shape in ONE launch (lockstep client
 * groups, see fc_gemm_bf16_grouped): every tensor argument becomes a host array of `groups` pointers. */
#define FC_ATTN_MAX_GROUPS 4
int fc_attention_fwd_grouped(int groups, const void* const* qkv, void* const* out, float* const* lse, int B, int N,
                             int H, int head_dim, int device, void* stream);
int fc_attention_bwd_grouped(int groups, const void* const* qkv, const void* const* out, const void* const* d_out,
                             const float* const* lse, void* const* dqkv, float* const* dbias, int B, int N, int H,
                             int head_dim, int device, void* stream);

/* ------------------------------------------------------------------------------------------------
 * LayerNorm over the fp32 residual stream           ref: src/models/mome.py:203,215,226-227,751-752
 *   fwd: y = LN(x) as bf16 (GEMM operand) and/or fp32; mean/rstd saved per row.
 *   bwd: dx (+)= LN'(dy); optional bf16 copy dxs = row_scale[row/rows_per_group] * dx for the next
 *        backward GEMM; dgamma/dbeta accumulated with atomics; dxs_colsum (nullable) += column sums of dxs
 *        (the bias gradient of the Linear that consumes dxs).
 * ------------------------------------------------------------------------------------------------ */
int fc_layernorm_fwd(const float* x, long long x_row_stride, const float* gamma, const float* beta, float eps,
                     void* y_bf16, float* y_f32, float* mean, float* rstd, int rows, int d, int device,
                     void* stream);
int fc_layernorm_bwd(const void* dy, int dy_is_bf16, long long dy_row_stride, const float* x,
                     long long x_row_stride, const float* mean, const float* rstd, const float* gamma,
                     float* dx, long long dx_row_stride, int accumulate, void* dxs_bf16,
                     long long dxs_row_stride, const float* row_scale, int rows_per_group, float* dgamma,
                     float* dbeta, float* dxs_colsum, int rows, int d, int device, void* stream);

/* Grouped forms (lockstep client groups, see fc_gemm_bf16_grouped): every tensor argument becomes a host array of
 * `groups` (<= FC_LN_MAX_GROUPS) pointers; shapes / strides are shared. */
#define FC_LN_MAX_GROUPS 4
int fc_layernorm_fwd_grouped(int groups, const float* const* x, long long x_row_stride, const float* const* gamma,
                             const float* const* beta, float eps, void* const* y_bf16, float* const* y_f32,
                             float* const* mean, float* const* rstd, int rows, int d, int device, void* stream);
int fc_layernorm_bwd_grouped(int groups, const void* const* dy, int dy_is_bf16, long long dy_row_stride,
                             const float* const* x, long long x_row_stride, const float* const* mean,
                             const float* const* rstd, const float* const* gamma, float* const* dx,
                             long long dx_row_stride, int accumulate, void* const* dxs_bf16, long long dxs_row_stride,
                             const float* const* row_scale, int rows_per_group, float* const* dgamma,
                             float* const* dbeta, float* const* dxs_colsum, int rows, int d, int device, void* stream);

/* ------------------------------------------------------------------------------------------------
 * Embeddings / heads / losses        ref: src/models/mome.py:597-611 (image), :632-639 (text, BertEmbeddings),
 *   :641-659 (heads), src/client/fedavgclient.py:85-95 (CrossEntropyLoss / ContrastiveLoss)
 * ------------------------------------------------------------------------------------------------ */
int fc_im2col16(const float* img, void* patches_bf16, float* x, const float* cls_token, const float* pos_embed,
                int B, int in_chans, int img_size, int d, int device, void* stream);
int fc_patch_bwd_prep(const float* dx, void* dxp_bf16, float* dpos, float* dcls, float* dbias, int B, int patches,
                      int d, int device, void* stream);
int fc_text_embed_fwd(const long long* ids, const float* word, const float* pos, const float* type,
                      const float* gamma, const float* beta, float eps, float* x, float* mean, float* rstd, int B,
                      int L, int d, int device, void* stream);
int fc_text_embed_bwd(const float* dx, const long long* ids, const float* word, const float* pos, const float* type,
                      const float* gamma, const float* mean, const float* rstd, float* dword, float* dpos,
                      float* dtype, float* dgamma, float* dbeta, int B, int L, int d, int device, void* stream);
int fc_head_fwd(const float* feat, const float* W, const float* bias, float* logits, int B, int d, int C,
                int device, void* stream);
int fc_head_bwd(const float* dlogits, const float* feat, const float* W, float* dW, float* dbias, float* dfeat,
                int B, int d, int C, int device, void* stream);
/* loss_out[0] += mean CE; dlogits = d(mean CE)/dlogits * grad_scale; correct_out[0] += #(argmax == target);
 * loss_sum_out[0] += sum of per-sample losses (what MetricManager.track accumulates: loss * len(batch)) */
int fc_ce_loss(const float* logits, const long long* target, float* dlogits, float* loss_out, float* correct_out,
               float* loss_sum_out, int B, int C, float grad_scale, int device, void* stream);
int fc_l2norm_fwd(const float* v, float* out, float* norm, int B, int d, int device, void* stream);
int fc_l2norm_bwd(const float* dout, const float* out, const float* norm, float* dv, int B, int d, int device,
                  void* stream);
/* sim_ws: fp32 [B*B] workspace, lse_ws: fp32 [2*B] workspace */
int fc_contrastive_loss(const float* a, const float* b, float* sim_ws, float* lse_ws, float* da, float* db,
                        float* loss_out, float* loss_sum_out, int B, int d, float tau, float grad_scale, int device,
                        void* stream);

/* ------------------------------------------------------------------------------------------------
 * Client optimizer step and friends     ref: src/client/fedavgclient.py:63,97-100 (AdamW/SGD, clip),
 *   src/client/fedproxclient.py:64-67 (prox term), src/models/mome.py:58-60 (aux mix)
 *   `chunks`: device array of {long long offset; int len; int segment;} work items (len <= fc_chunk_floats()).
 *   grad_sumsq/max_norm: when max_norm > 0 the gradient is scaled by min(1, max_norm/(sqrt(*grad_sumsq)+1e-6)).
 * ------------------------------------------------------------------------------------------------ */
int fc_chunk_floats(void);
int fc_adamw_step(float* params, const float* grads, float* exp_avg, float* exp_avg_sq, const void* chunks,
                  int n_chunks, float lr, float beta1, float beta2, float eps, float weight_decay, int step,
                  const float* grad_sumsq, float max_norm, int device, void* stream);
int fc_sgd_step(float* params, const float* grads, float* momentum_buf, const void* chunks, int n_chunks, float lr,
                float momentum, float dampening, float weight_decay, int nesterov, int first_step,
                const float* grad_sumsq, float max_norm, int device, void* stream);
/* out[segment] (or out[0] if single_output) += sum (a-b)^2 ; b may be NULL */
int fc_sumsq(const float* a, const float* b, const void* chunks, int n_chunks, float* out, int single_output,
             int device, void* stream);
/* grads += mu*0.5*(p - pg)/||p - pg||_segment ; loss_out[0] += L, loss_weighted_out[0] += weight*L,
 * L = mu*0.5*sum_segments ||p - pg|| */
int fc_prox_grad(float* grads, const float* params, const float* global_params, const void* chunks, int n_chunks,
                 const float* seg_sumsq, int n_segments, float mu, float* loss_out, float* loss_weighted_out,
                 float weight, int device, void* stream);
/* layers: device array of {long long w_off, a_off, s_off, dst_off, dstT_off; int rows, cols, tile_start;} */
int fc_prep_weights(const float* params, void* operands_bf16, const void* layers, int n_layers, int n_tiles,
                    int device, void* stream);
/* layers: device array of {long long w_off, a_off, s_off, numel; int chunk_start; int pad;} */
int fc_aux_grads(const float* params, float* grads, const void* layers, int n_layers, int n_chunks,
                 int aux_trained, int device, void* stream);
int fc_colsum_bf16(const void* x, long long ld, int rows, int n, float* out, int device, void* stream);
/* The optimizer tail in two launches instead of aux_grads -> step -> prep_weights (no clipping, no FedProx term):
 * phase A steps every trainable tensor that is not a Linear weight — an aux_weight chunk takes its gradient s_old*dW on
 * the fly and leaves its share of ds = <dW, A>; the block finishing a layer steps its cross_modal_scale — and phase B
 * steps the Linear weights and writes the bf16 GEMM operand bf16(W_new + s_new*A_new) from registers.
 *   chunks_a / chunks_b: device arrays of {long long off; int len, kind, layer, pad; long long x0, x1;}
 *     kind 0 plain | 1 aux_weight (x0 = offset of the matching W element) | 2 Linear weight (x0 = operand element)
 *     | 3 Linear weight with aux partner (x0 = operand element, x1 = offset of the matching aux_weight element);
 *   aux_layers as for fc_aux_grads; counters: one zeroed int per aux layer (zero again on return).
 * ref: src/models/mome.py:58-60 + autograd, src/client/fedavgclient.py:63,100. */
int fc_opt_fused(float* params, float* grads, float* state0, float* state1, void* operands_bf16, const void* chunks_a,
                 int n_chunks_a, const void* chunks_b, int n_chunks_b, const void* aux_layers, int* counters,
                 int optimizer, float lr, float beta1, float beta2, float eps, float weight_decay, float momentum,
                 float dampening, int nesterov, int step, int device, void* stream);

/* ------------------------------------------------------------------------------------------------
 * fp32-accurate validation mode (precision = 'fp32'): fp32 activations, split-operand GEMMs (fc_gemm_split), fp32 FMA
 * attention.  ref: the reference's strict-fp32 arithmetic, src/models/mome.py:150-168.  Simple kernels, not tuned.
 * ------------------------------------------------------------------------------------------------ */
/* x (fp32, n elements, rows of row_len) [* row_scale[row / rows_per_group]] -> hi = bf16(x), lo = bf16(x - hi) */
int fc_split_bf16(const float* x, void* hi, void* lo, long long n, int row_len, const float* row_scale,
                  int rows_per_group, int device, void* stream);
/* W_eff = W + s*A of every Linear as (hi, lo) pairs; `layers` as for fc_prep_weights */
int fc_prep_weights_split(const float* params, void* hi, void* lo, const void* layers, int n_layers, int device,
                          void* stream);
int fc_gelu_f32_fwd(const float* pre, float* act, long long n, int device, void* stream);
int fc_gelu_f32_bwd(const float* d_act, const float* pre, float* d_pre, long long n, int device, void* stream);
/* out[n] += column sums of x[rows, n] (optionally row-scaled) */
int fc_colsum_f32(const float* x, long long ld, int rows, int n, const float* row_scale, int rows_per_group, float* out,
                  int device, void* stream);
/* qkv fp32 [B, N, 3, H, 64]; out fp32 [B, N, H*64]; lse fp32 [B, H, N]; dqkv must be zeroed by the caller */
int fc_attention_f32_fwd(const float* qkv, float* out, float* lse, int B, int N, int H, int device, void* stream);
int fc_attention_f32_bwd(const float* qkv, const float* out, const float* d_out, const float* lse, float* dqkv, int B,
                         int N, int H, int device, void* stream);
int fc_im2col16_f32(const float* img, float* patches, int B, int in_chans, int img_size, int device, void* stream);
int fc_drop_cls_rows(const float* dx, float* dxp, int B, int patches, int d, int device, void* stream);

/* ------------------------------------------------------------------------------------------------
 * Native step driver                 ref: src/models/mome.py:881-922 (ModalityAgnosticTransformer.forward),
 *   src/client/fedavgclient.py:79-102 and src/client/fedproxclient.py:64-71 (one batch of the update loop)
 * ------------------------------------------------------------------------------------------------ */
#define FC_MAX_DEPTH 24
#define FC_MAX_SEGMENTS 2048
#define FC_BLOCK_ROLES 20   /* n1w n1b qkvw qkvb qkvs qkva projw projb projs proja n2w n2b fc1w fc1b fc1s fc1a fc2w fc2b fc2s fc2a */

typedef struct {
  int d, depth, heads, hidden;          /* embed dim, #blocks, #heads (head_dim 64), mlp hidden */
  int img_size, patches, in_chans;      /* 224, 196, 3|1 */
  int seq_len, vocab, max_text_len;
  int num_classes[2];                   /* per encoder slot (0 = img, 1 = txt); <= 0: retrieval / no head */
  int has_enc[2];
  int with_aux, aux_trained;
  int precise;                          /* 1: fp32-accurate validation mode (split-operand GEMMs, fp32 activations); the
                                           operand arena then holds W_eff hi parts followed, op_lo_offset elements later,
                                           by the lo parts */
  long long op_lo_offset;
  /* float offsets into the flat param arena (the grad / optimizer arenas share the layout); -1 = absent */
  long long img_pos, img_cls, img_pw, img_pb;
  long long txt_word, txt_pos, txt_type, txt_lnw, txt_lnb;
  long long norm_w, norm_b;
  long long head_w[2], head_b[2];
  long long blk[2][FC_MAX_DEPTH][FC_BLOCK_ROLES];
  /* bf16 element offsets into the operand arena: W_eff = W + s*A of every Linear, [out, in] row-major */
  long long op_pw;
  long long op[2][FC_MAX_DEPTH][4];     /* qkv, proj, fc1, fc2 */
} fc_mat_desc;

long long fc_mat_workspace_bytes(const fc_mat_desc* m, int B);
/* out0/out1: logits fp32 [B, C] or (feat_out / retrieval) unit features fp32 [B, d]; NULL to skip the copy.
 * droppath: fp32 [2 enc][depth][2 branches][B] per-sample scales (keep mask / keep prob) or NULL. */
int fc_mat_forward(const fc_mat_desc* m, const float* params, const void* operands, void* workspace, int B,
                   const float* img, const long long* ids, const float* droppath, int feat_out, float* out0,
                   float* out1, int device, void* stream);
/* Accumulates into grads (caller zeroes). dout0/dout1: gradient w.r.t. the forward outputs (NULL = none). */
int fc_mat_backward(const fc_mat_desc* m, const float* params, const void* operands, void* workspace, int B,
                    const long long* ids, const float* droppath, int feat_out, const float* dout0,
                    const float* dout1, float* grads, const void* aux_layers, int n_aux_layers,
                    int n_aux_chunks, int device, void* stream);

#define FC_LOSS_CE_IMG 0
#define FC_LOSS_CE_TXT 1
#define FC_LOSS_CONTRASTIVE 2
#define FC_OPT_NONE 0     /* forward + loss + backward only (gradients left in `grads`) */
#define FC_OPT_ADAMW 1
#define FC_OPT_SGD 2

typedef struct {
  int B;
  int loss_kind, optimizer, step;       /* step: 1-based count since the optimizer was created */
  float lr, beta1, beta2, eps, weight_decay, momentum, dampening;
  int nesterov;
  float max_grad_norm, prox_mu;
  float* params;                        /* flat fp32 arena, updated in place */
  float* grads;                         /* same layout; zeroed by the call */
  float* opt_state0;                    /* AdamW exp_avg / SGD momentum buffer */
  float* opt_state1;                    /* AdamW exp_avg_sq */
  const float* global_params;           /* FedProx: frozen copy of the downloaded model */
  void* operands;                       /* bf16 operand arena */
  void* workspace;
  long long arena_floats;
  const float* img;                     /* fp32 [B, C, 224, 224] */
  const long long* ids;                 /* int64 [B, seq_len] */
  const long long* labels;              /* int64 [B] (CE) */
  const float* droppath;
  float* stats;                         /* [3]: += step loss (incl. prox term), += #correct (CE only),
                                           += step loss * B (MetricManager's running loss) */
  const void* chunks; int n_chunks; int n_segments;      /* trainable chunk table */
  const void* prep_layers; int n_prep_layers; int n_prep_tiles;
  const void* aux_layers; int n_aux_layers; int n_aux_chunks;
  /* fused optimizer tail (fc_opt_fused); n_fused_a + n_fused_b == 0: the separate aux_grads / step / prep kernels */
  const void* fused_a; int n_fused_a;
  const void* fused_b; int n_fused_b;
  int* fused_counters;
} fc_step_args;

int fc_client_step(const fc_mat_desc* m, const fc_step_args* a, int device, void* stream);
/* The same step for a lockstep group of `n` (<= FC_STEP_MAX_GROUPS) clients of ONE model architecture at one batch size:
 * every GEMM / attention / LayerNorm launch of the step covers all of them (fc_*_grouped), so the fixed cost of a
 * launch — prologue, first-load latency, exposed last epilogue, wave quantisation: 40-50 % of a ViT-S GEMM at B=112 —
 * is paid once per group.  ref: the clients the reference's ThreadPoolExecutor trains side by side
 * (src/server/fedavgserver.py:566-577), each running src/client/fedavgclient.py:79-102. */
#define FC_STEP_MAX_GROUPS 4
int fc_client_step_group(const fc_mat_desc* m, int n, const fc_step_args* const* args, int device, void* stream);

/* struct-layout handshake for FFI bindings */
int fc_sizeof_mat_desc(void);
int fc_sizeof_step_args(void);

#ifdef __cplusplus
}
#endif
#endif /* FEDCOLA_B200_H_ */
