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

const char* fc_last_error(void);
int fc_abi_version(void);

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
#define FC_EPI_GELU 1        /* out(bf16) = acc + bias ; out2(bf16) = gelu_erf(acc + bias)          */
#define FC_EPI_RESID 2       /* out(f32)  = resid + row_scale[row/rows_per_group] * (acc + bias)    */
#define FC_EPI_DGELU 3       /* out(bf16) = acc * gelu_erf'(aux)     (aux = bf16 pre-activation)    */
#define FC_EPI_F32 4         /* out(f32)  = acc + bias                                              */
#define FC_EPI_ATOMIC_F32 5  /* out(f32) += alpha * acc  (red.add; split-K over blockIdx.z)         */
#define FC_EPI_PATCH 6       /* out(f32)[b*(P+1)+1+t] = acc + bias + pos[1+t]  (row = b*P + t)      */

int fc_gemm_bf16(int M, int N, int K, const void* A, long long lda, int a_mn_major, const void* B,
                 long long ldb, int b_mn_major, int epi, void* out, void* out2, long long ldo,
                 const float* bias, const float* resid, const float* row_scale, int rows_per_group,
                 const void* aux, const float* pos, int patches, float alpha, int splits, int device,
                 void* stream);

#ifdef __cplusplus
}
#endif
#endif /* FEDCOLA_B200_H_ */
