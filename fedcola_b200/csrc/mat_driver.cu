// Native step driver: one C call runs the whole ModalityAgnosticTransformer forward, or backward, or a full
// client training step (forward + loss + backward + clip/prox + optimizer + operand refresh) on one stream.
//
// Replaces the Python-level op-by-op execution of ModalityAgnosticTransformer.forward
// (/root/reference/src/models/mome.py:881-922; ~8 400 aten calls per ViT-S step) and the body of
// the batch loop of FedavgClient.update / FedproxClient.update
// (/root/reference/src/client/fedavgclient.py:79-102, fedproxclient.py:64-71).
//
// All memory is caller-owned: the flat fp32 param/grad/optimizer arenas, the bf16 operand arena, and one
// activation workspace whose size fc_mat_workspace_bytes() reports.  No global state; re-entrant.
#include "common.cuh"
#include "../../include/fedcola_b200.h"

namespace {

enum Role { N1W, N1B, QKVW, QKVB, QKVS, QKVA, PROJW, PROJB, PROJS, PROJA, N2W, N2B, FC1W, FC1B, FC1S, FC1A, FC2W, FC2B,
            FC2S, FC2A };
enum Lin { L_QKV, L_PROJ, L_FC1, L_FC2 };

inline size_t align_up(size_t x, size_t a = 256) { return (x + a - 1) / a * a; }

// Carves the activation workspace. Every pointer is 256-byte aligned.
struct EncWs {
  float* x_in[FC_MAX_DEPTH + 1];
  float* x_mid[FC_MAX_DEPTH];
  __nv_bfloat16 *ln1[FC_MAX_DEPTH], *ln2[FC_MAX_DEPTH], *qkv[FC_MAX_DEPTH], *ao[FC_MAX_DEPTH], *hgrad[FC_MAX_DEPTH],
      *hact[FC_MAX_DEPTH];
  float *mean1[FC_MAX_DEPTH], *rstd1[FC_MAX_DEPTH], *mean2[FC_MAX_DEPTH], *rstd2[FC_MAX_DEPTH], *lse[FC_MAX_DEPTH];
  __nv_bfloat16* patches;     // img: [B*P, 768]
  float *mean_e, *rstd_e;     // txt embedding LN
  float *mean_f, *rstd_f, *feat, *featn, *fnorm, *logits;
  // backward temporaries
  float* dx;
  __nv_bfloat16 *dxs, *d_h, *d_ln, *d_ao, *d_qkv, *dxp;
  float *dfeat, *dout_tmp;
};
struct Ws {
  EncWs enc[2];
  float *sim, *lse_ws, *da, *db, *dlogits;     // loss scratch
  float* scalars;                              // [8]: 0 grad sumsq, 1.. spare
  float* seg_sumsq;                            // [max segments] FedProx per-tensor norms
  size_t bytes;
};

struct Carver {
  uint8_t* base;
  size_t off = 0;
  template <typename T>
  T* take(size_t n) {
    T* p = base ? reinterpret_cast<T*>(base + off) : nullptr;
    off = align_up(off + n * sizeof(T));
    return p;
  }
};

int tokens_of(const fc_mat_desc* m, int e) { return e == 0 ? m->patches + 1 : m->seq_len; }

void carve(const fc_mat_desc* m, int B, void* base, Ws* w) {
  Carver c{reinterpret_cast<uint8_t*>(base)};
  const int d = m->d, hid = m->hidden, L = m->depth, H = m->heads;
  int maxC = 1;
  for (int e = 0; e < 2; ++e) {
    if (!m->has_enc[e]) continue;
    EncWs& s = w->enc[e];
    const size_t N = tokens_of(m, e), T = (size_t)B * N;
    for (int j = 0; j <= L; ++j) s.x_in[j] = c.take<float>(T * d);
    for (int j = 0; j < L; ++j) {
      s.x_mid[j] = c.take<float>(T * d);
      s.ln1[j] = c.take<__nv_bfloat16>(T * d);
      s.ln2[j] = c.take<__nv_bfloat16>(T * d);
      s.qkv[j] = c.take<__nv_bfloat16>(T * 3 * d);
      s.ao[j] = c.take<__nv_bfloat16>(T * d);
      s.hgrad[j] = c.take<__nv_bfloat16>(T * hid);
      s.hact[j] = c.take<__nv_bfloat16>(T * hid);
      s.mean1[j] = c.take<float>(T);
      s.rstd1[j] = c.take<float>(T);
      s.mean2[j] = c.take<float>(T);
      s.rstd2[j] = c.take<float>(T);
      s.lse[j] = c.take<float>((size_t)B * H * N);
    }
    s.patches = e == 0 ? c.take<__nv_bfloat16>((size_t)B * m->patches * 768) : nullptr;
    s.mean_e = c.take<float>(T);
    s.rstd_e = c.take<float>(T);
    s.mean_f = c.take<float>(B);
    s.rstd_f = c.take<float>(B);
    s.feat = c.take<float>((size_t)B * d);
    s.featn = c.take<float>((size_t)B * d);
    s.fnorm = c.take<float>(B);
    const int C = m->num_classes[e] > 0 ? m->num_classes[e] : 1;
    if (C > maxC) maxC = C;
    s.logits = c.take<float>((size_t)B * C);
    s.dx = c.take<float>(T * d);
    s.dxs = c.take<__nv_bfloat16>(T * d);
    s.d_h = c.take<__nv_bfloat16>(T * hid);
    s.d_ln = c.take<__nv_bfloat16>(T * d);
    s.d_ao = c.take<__nv_bfloat16>(T * d);
    s.d_qkv = c.take<__nv_bfloat16>(T * 3 * d);
    s.dxp = e == 0 ? c.take<__nv_bfloat16>((size_t)B * m->patches * d) : nullptr;
    s.dfeat = c.take<float>((size_t)B * d);
    s.dout_tmp = c.take<float>((size_t)B * (d > C ? d : C));
  }
  w->sim = c.take<float>((size_t)B * B);
  w->lse_ws = c.take<float>(2 * (size_t)B);
  w->da = c.take<float>((size_t)B * d);
  w->db = c.take<float>((size_t)B * d);
  w->dlogits = c.take<float>((size_t)B * maxC);
  w->scalars = c.take<float>(8);
  w->seg_sumsq = c.take<float>(FC_MAX_SEGMENTS);
  w->bytes = c.off;
}

#define TRY(expr)            \
  do {                       \
    int _rc = (expr);        \
    if (_rc != FC_OK) return _rc; \
  } while (0)

struct Ctx {
  const fc_mat_desc* m;
  const float* params;
  float* grads;
  const __nv_bfloat16* ops;
  int B, device;
  void* stream;
  const float* droppath;    // [2 enc][L][2][B] or null
  const float* p(long long off) const { return off >= 0 ? params + off : nullptr; }
  float* g(long long off) const { return off >= 0 ? grads + off : nullptr; }
  const float* dp(int e, int j, int which) const {
    return droppath ? droppath + (((size_t)e * m->depth + j) * 2 + which) * B : nullptr;
  }
};

int gemm(const Ctx& c, int M, int N, int K, const void* A, long long lda, int a_mn, const void* Bm, long long ldb,
         int b_mn, int epi, void* out, void* out2, long long ldo, const float* bias, const float* resid,
         const float* row_scale, int rpg, const void* aux, const float* pos, int patches, int splits,
         float* colsum = nullptr) {
  return fc_gemm_bf16(M, N, K, A, lda, a_mn, Bm, ldb, b_mn, epi, out, out2, ldo, bias, resid, row_scale, rpg, aux, pos,
                      patches, 1.0f, splits, colsum, c.device, c.stream);
}

// dW[rows_out, cols_out] += dY[T, rows_out]^T X[T, cols_out]   (both operands MN-major, split-K over tokens)
int gemm_dw(const Ctx& c, int rows_out, int cols_out, int T, const void* dY, const void* X, float* dW) {
  return gemm(c, rows_out, cols_out, T, dY, rows_out, 1, X, cols_out, 1, FC_EPI_ATOMIC_F32, dW, nullptr, cols_out,
              nullptr, nullptr, nullptr, 0, nullptr, nullptr, 0, /*splits: auto*/ 0);
}

int encoder_forward(const Ctx& c, Ws& w, int e, const float* img, const long long* ids) {
  const fc_mat_desc* m = c.m;
  EncWs& s = w.enc[e];
  const int d = m->d, hid = m->hidden, L = m->depth, H = m->heads, B = c.B;
  const int N = tokens_of(m, e), T = B * N;
  if (e == 0) {
    FC_REQUIRE(img != nullptr, "fc_mat_forward: image input missing");
    TRY(fc_im2col16(img, s.patches, s.x_in[0], c.p(m->img_cls), c.p(m->img_pos), B, m->in_chans, m->img_size, d,
                    c.device, c.stream));
    TRY(gemm(c, B * m->patches, d, 768, s.patches, 768, 0, c.ops + m->op_pw, 768, 0, FC_EPI_PATCH, s.x_in[0], nullptr,
             d, c.p(m->img_pb), nullptr, nullptr, 0, nullptr, c.p(m->img_pos), m->patches, 1));
  } else {
    FC_REQUIRE(ids != nullptr, "fc_mat_forward: token ids missing");
    TRY(fc_text_embed_fwd(ids, c.p(m->txt_word), c.p(m->txt_pos), c.p(m->txt_type), c.p(m->txt_lnw), c.p(m->txt_lnb),
                          1e-12f, s.x_in[0], s.mean_e, s.rstd_e, B, N, d, c.device, c.stream));
  }
  for (int j = 0; j < L; ++j) {
    const long long* o = m->blk[e][j];
    const long long* op = m->op[e][j];
    TRY(fc_layernorm_fwd(s.x_in[j], d, c.p(o[N1W]), c.p(o[N1B]), 1e-5f, s.ln1[j], nullptr, s.mean1[j], s.rstd1[j], T, d,
                         c.device, c.stream));
    TRY(gemm(c, T, 3 * d, d, s.ln1[j], d, 0, c.ops + op[L_QKV], d, 0, FC_EPI_BF16, s.qkv[j], nullptr, 3 * d,
             c.p(o[QKVB]), nullptr, nullptr, 0, nullptr, nullptr, 0, 1));
    TRY(fc_attention_fwd(s.qkv[j], s.ao[j], s.lse[j], B, N, H, d / H, c.device, c.stream));
    TRY(gemm(c, T, d, d, s.ao[j], d, 0, c.ops + op[L_PROJ], d, 0, FC_EPI_RESID, s.x_mid[j], nullptr, d, c.p(o[PROJB]),
             s.x_in[j], c.dp(e, j, 0), N, nullptr, nullptr, 0, 1));
    TRY(fc_layernorm_fwd(s.x_mid[j], d, c.p(o[N2W]), c.p(o[N2B]), 1e-5f, s.ln2[j], nullptr, s.mean2[j], s.rstd2[j], T,
                         d, c.device, c.stream));
    TRY(gemm(c, T, hid, d, s.ln2[j], d, 0, c.ops + op[L_FC1], d, 0, FC_EPI_GELU, s.hgrad[j], s.hact[j], hid,
             c.p(o[FC1B]), nullptr, nullptr, 0, nullptr, nullptr, 0, 1));
    TRY(gemm(c, T, d, hid, s.hact[j], hid, 0, c.ops + op[L_FC2], hid, 0, FC_EPI_RESID, s.x_in[j + 1], nullptr, d,
             c.p(o[FC2B]), s.x_mid[j], c.dp(e, j, 1), N, nullptr, nullptr, 0, 1));
  }
  // final norm (eps 1e-6) — only the cls token is consumed by the heads (mome.py:647,658,915)
  TRY(fc_layernorm_fwd(s.x_in[L], (long long)N * d, c.p(m->norm_w), c.p(m->norm_b), 1e-6f, nullptr, s.feat, s.mean_f,
                       s.rstd_f, B, d, c.device, c.stream));
  return FC_OK;
}

// dfeat (fp32 [B,d], gradient w.r.t. the final-norm'ed cls token) is in s.dfeat
int encoder_backward(const Ctx& c, Ws& w, int e, const long long* ids) {
  const fc_mat_desc* m = c.m;
  EncWs& s = w.enc[e];
  const int d = m->d, hid = m->hidden, L = m->depth, H = m->heads, B = c.B;
  const int N = tokens_of(m, e), T = B * N;
  cudaStream_t st = reinterpret_cast<cudaStream_t>(c.stream);
  FC_CUDA_CHECK(cudaMemsetAsync(s.dx, 0, sizeof(float) * (size_t)T * d, st));
  FC_CUDA_CHECK(cudaMemsetAsync(s.dxs, 0, sizeof(__nv_bfloat16) * (size_t)T * d, st));
  // final norm backward on the cls rows; dxs = DropPath scale of the last block's mlp branch * dx
  TRY(fc_layernorm_bwd(s.dfeat, 0, d, s.x_in[L], (long long)N * d, s.mean_f, s.rstd_f, c.p(m->norm_w), s.dx,
                       (long long)N * d, 0, s.dxs, (long long)N * d, c.dp(e, L - 1, 1), 1, c.g(m->norm_w),
                       c.g(m->norm_b), c.g(m->blk[e][L - 1][FC2B]), B, d, c.device, c.stream));
  for (int j = L - 1; j >= 0; --j) {
    const long long* o = m->blk[e][j];
    const long long* op = m->op[e][j];
    // ---- mlp branch:  x_out = x_mid + dp2 * (fc2(gelu(fc1(LN2(x_mid)))))
    // (fc2 bias gradient = column sums of dxs: accumulated by the LayerNorm backward that produced dxs)
    TRY(gemm(c, T, hid, d, s.dxs, d, 0, c.ops + op[L_FC2], hid, 1, FC_EPI_MULAUX, s.d_h, nullptr, hid, nullptr, nullptr,
             nullptr, 0, s.hgrad[j], nullptr, 0, 1, c.g(o[FC1B])));     // d_h = (dxs W2) * gelu'(pre); fc1 bias grad fused
    TRY(gemm_dw(c, d, hid, T, s.dxs, s.hact[j], c.g(o[FC2W])));
    TRY(gemm(c, T, d, hid, s.d_h, hid, 0, c.ops + op[L_FC1], d, 1, FC_EPI_BF16, s.d_ln, nullptr, d, nullptr, nullptr,
             nullptr, 0, nullptr, nullptr, 0, 1));
    TRY(gemm_dw(c, hid, d, T, s.d_h, s.ln2[j], c.g(o[FC1W])));
    TRY(fc_layernorm_bwd(s.d_ln, 1, d, s.x_mid[j], d, s.mean2[j], s.rstd2[j], c.p(o[N2W]), s.dx, d, 1, s.dxs, d,
                         c.dp(e, j, 0), N, c.g(o[N2W]), c.g(o[N2B]), c.g(o[PROJB]), T, d, c.device, c.stream));
    // ---- attention branch:  x_mid = x_in + dp1 * proj(attn(qkv(LN1(x_in))))
    TRY(gemm(c, T, d, d, s.dxs, d, 0, c.ops + op[L_PROJ], d, 1, FC_EPI_BF16, s.d_ao, nullptr, d, nullptr, nullptr,
             nullptr, 0, nullptr, nullptr, 0, 1));
    TRY(gemm_dw(c, d, d, T, s.dxs, s.ao[j], c.g(o[PROJW])));
    TRY(fc_attention_bwd(s.qkv[j], s.ao[j], s.d_ao, s.lse[j], s.d_qkv, c.g(o[QKVB]), B, N, H, d / H, c.device,
                         c.stream));
    TRY(gemm(c, T, d, 3 * d, s.d_qkv, 3 * d, 0, c.ops + op[L_QKV], d, 1, FC_EPI_BF16, s.d_ln, nullptr, d, nullptr,
             nullptr, nullptr, 0, nullptr, nullptr, 0, 1));
    TRY(gemm_dw(c, 3 * d, d, T, s.d_qkv, s.ln1[j], c.g(o[QKVW])));
    TRY(fc_layernorm_bwd(s.d_ln, 1, d, s.x_in[j], d, s.mean1[j], s.rstd1[j], c.p(o[N1W]), s.dx, d, 1, s.dxs, d,
                         j > 0 ? c.dp(e, j - 1, 1) : nullptr, N, c.g(o[N1W]), c.g(o[N1B]),
                         j > 0 ? c.g(m->blk[e][j - 1][FC2B]) : nullptr, T, d, c.device, c.stream));
  }
  if (e == 0) {
    TRY(fc_patch_bwd_prep(s.dx, s.dxp, c.g(m->img_pos), c.g(m->img_cls), c.g(m->img_pb), B, m->patches, d, c.device,
                          c.stream));
    TRY(gemm_dw(c, d, 768, B * m->patches, s.dxp, s.patches, c.g(m->img_pw)));
  } else {
    TRY(fc_text_embed_bwd(s.dx, ids, c.p(m->txt_word), c.p(m->txt_pos), c.p(m->txt_type), c.p(m->txt_lnw), s.mean_e,
                          s.rstd_e, c.g(m->txt_word), c.g(m->txt_pos), c.g(m->txt_type), c.g(m->txt_lnw),
                          c.g(m->txt_lnb), B, N, d, c.device, c.stream));
  }
  return FC_OK;
}

bool is_retrieval(const fc_mat_desc* m, int e, int feat_out) { return feat_out || m->num_classes[e] <= 0; }

int check_desc(const fc_mat_desc* m, int B) {
  FC_REQUIRE(m != nullptr, "null model descriptor");
  FC_REQUIRE(m->depth >= 1 && m->depth <= FC_MAX_DEPTH, "depth %d out of range", m->depth);
  FC_REQUIRE(m->d % 64 == 0 && m->heads * 64 == m->d, "embed_dim must be heads*64 (head_dim 64)");
  FC_REQUIRE(m->hidden % 8 == 0, "mlp hidden size must be a multiple of 8");
  FC_REQUIRE(m->has_enc[0] || m->has_enc[1], "model has no encoder");
  FC_REQUIRE(!m->has_enc[0] || (m->img_size == 224 && m->patches == 196 && (m->in_chans == 3 || m->in_chans == 1)),
             "image encoder must be 224x224 with 16x16 patches");
  FC_REQUIRE(!m->has_enc[1] || (m->seq_len >= 1 && m->seq_len <= 256), "seq_len must be in [1,256]");
  FC_REQUIRE(B >= 1, "batch must be >= 1");
  return FC_OK;
}

}  // namespace

extern "C" long long fc_mat_workspace_bytes(const fc_mat_desc* m, int B) {
  if (check_desc(m, B) != FC_OK) return -1;
  Ws w;
  carve(m, B, nullptr, &w);
  return (long long)w.bytes;
}

extern "C" int fc_mat_forward(const fc_mat_desc* m, const float* params, const void* operands, void* workspace, int B,
                              const float* img, const long long* ids, const float* droppath, int feat_out,
                              float* out0, float* out1, int device, void* stream) {
  TRY(check_desc(m, B));
  FcDeviceGuard guard(device);
  Ws w;
  carve(m, B, workspace, &w);
  Ctx c{m, params, nullptr, reinterpret_cast<const __nv_bfloat16*>(operands), B, device, stream, droppath};
  float* outs[2] = {out0, out1};
  for (int e = 0; e < 2; ++e) {
    if (!m->has_enc[e]) continue;
    TRY(encoder_forward(c, w, e, img, ids));
    EncWs& s = w.enc[e];
    if (is_retrieval(m, e, feat_out)) {
      TRY(fc_l2norm_fwd(s.feat, s.featn, s.fnorm, B, m->d, device, stream));
      if (outs[e])
        FC_CUDA_CHECK(cudaMemcpyAsync(outs[e], s.featn, sizeof(float) * (size_t)B * m->d, cudaMemcpyDeviceToDevice,
                                      reinterpret_cast<cudaStream_t>(stream)));
    } else {
      TRY(fc_head_fwd(s.feat, c.p(m->head_w[e]), c.p(m->head_b[e]), s.logits, B, m->d, m->num_classes[e], device,
                      stream));
      if (outs[e])
        FC_CUDA_CHECK(cudaMemcpyAsync(outs[e], s.logits, sizeof(float) * (size_t)B * m->num_classes[e],
                                      cudaMemcpyDeviceToDevice, reinterpret_cast<cudaStream_t>(stream)));
    }
  }
  return FC_OK;
}

extern "C" int fc_mat_backward(const fc_mat_desc* m, const float* params, const void* operands, void* workspace,
                               int B, const long long* ids, const float* droppath, int feat_out, const float* dout0,
                               const float* dout1, float* grads, const void* aux_layers, int n_aux_layers,
                               int n_aux_chunks, int device, void* stream) {
  TRY(check_desc(m, B));
  FcDeviceGuard guard(device);
  Ws w;
  carve(m, B, workspace, &w);
  Ctx c{m, params, grads, reinterpret_cast<const __nv_bfloat16*>(operands), B, device, stream, droppath};
  const float* douts[2] = {dout0, dout1};
  for (int e = 0; e < 2; ++e) {
    if (!m->has_enc[e] || douts[e] == nullptr) continue;
    EncWs& s = w.enc[e];
    if (is_retrieval(m, e, feat_out)) {
      TRY(fc_l2norm_bwd(douts[e], s.featn, s.fnorm, s.dfeat, B, m->d, device, stream));
    } else {
      TRY(fc_head_bwd(douts[e], s.feat, c.p(m->head_w[e]), c.g(m->head_w[e]), c.g(m->head_b[e]), s.dfeat, B, m->d,
                      m->num_classes[e], device, stream));
    }
    TRY(encoder_backward(c, w, e, ids));
  }
  if (n_aux_layers > 0)
    TRY(fc_aux_grads(params, grads, aux_layers, n_aux_layers, n_aux_chunks, m->aux_trained, device, stream));
  return FC_OK;
}

extern "C" int fc_client_step(const fc_mat_desc* m, const fc_step_args* a, int device, void* stream) {
  TRY(check_desc(m, a ? a->B : 0));
  FC_REQUIRE(a->loss_kind >= FC_LOSS_CE_IMG && a->loss_kind <= FC_LOSS_CONTRASTIVE, "bad loss kind %d", a->loss_kind);
  FcDeviceGuard guard(device);
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  const int B = a->B, d = m->d;
  Ws w;
  carve(m, B, a->workspace, &w);
  FC_REQUIRE(a->n_segments <= FC_MAX_SEGMENTS, "too many parameter segments (%d)", a->n_segments);
  Ctx c{m, a->params, a->grads, reinterpret_cast<const __nv_bfloat16*>(a->operands), B, device, stream, a->droppath};

  // ---- forward
  for (int e = 0; e < 2; ++e)
    if (m->has_enc[e]) TRY(encoder_forward(c, w, e, a->img, a->ids));
  FC_CUDA_CHECK(cudaMemsetAsync(a->grads, 0, sizeof(float) * (size_t)a->arena_floats, st));   // optimizer.zero_grad()

  // ---- loss + its gradient w.r.t. the encoder outputs
  if (a->loss_kind == FC_LOSS_CONTRASTIVE) {
    FC_REQUIRE(m->has_enc[0] && m->has_enc[1], "contrastive loss needs both encoders");
    for (int e = 0; e < 2; ++e) TRY(fc_l2norm_fwd(w.enc[e].feat, w.enc[e].featn, w.enc[e].fnorm, B, d, device, stream));
    TRY(fc_contrastive_loss(w.enc[0].featn, w.enc[1].featn, w.sim, w.lse_ws, w.da, w.db, a->stats, a->stats + 2, B, d,
                            1.0f / 0.07f, 1.0f, device, stream));
    TRY(fc_l2norm_bwd(w.da, w.enc[0].featn, w.enc[0].fnorm, w.enc[0].dfeat, B, d, device, stream));
    TRY(fc_l2norm_bwd(w.db, w.enc[1].featn, w.enc[1].fnorm, w.enc[1].dfeat, B, d, device, stream));
    TRY(encoder_backward(c, w, 0, a->ids));
    TRY(encoder_backward(c, w, 1, a->ids));
  } else {
    const int e = a->loss_kind == FC_LOSS_CE_IMG ? 0 : 1;
    FC_REQUIRE(m->has_enc[e] && m->num_classes[e] > 0 && a->labels != nullptr, "CE loss needs a classification head and labels");
    EncWs& s = w.enc[e];
    const int C = m->num_classes[e];
    TRY(fc_head_fwd(s.feat, c.p(m->head_w[e]), c.p(m->head_b[e]), s.logits, B, d, C, device, stream));
    TRY(fc_ce_loss(s.logits, a->labels, w.dlogits, a->stats, a->stats + 1, a->stats + 2, B, C, 1.0f, device, stream));
    TRY(fc_head_bwd(w.dlogits, s.feat, c.p(m->head_w[e]), c.g(m->head_w[e]), c.g(m->head_b[e]), s.dfeat, B, d, C,
                    device, stream));
    TRY(encoder_backward(c, w, e, a->ids));
  }
  if (a->n_aux_layers > 0)
    TRY(fc_aux_grads(a->params, a->grads, a->aux_layers, a->n_aux_layers, a->n_aux_chunks, m->aux_trained, device, stream));

  // ---- FedProx proximal term (fedproxclient.py:64-67): per-tensor un-squared L2 norms
  if (a->prox_mu > 0.f && a->global_params != nullptr) {
    FC_CUDA_CHECK(cudaMemsetAsync(w.seg_sumsq, 0, sizeof(float) * a->n_segments, st));
    TRY(fc_sumsq(a->params, a->global_params, a->chunks, a->n_chunks, w.seg_sumsq, 0, device, stream));
    TRY(fc_prox_grad(a->grads, a->params, a->global_params, a->chunks, a->n_chunks, w.seg_sumsq, a->n_segments,
                     a->prox_mu, a->stats, a->stats + 2, (float)B, device, stream));
  }
  // ---- clip_grad_norm_ (fedavgclient.py:98-99): global L2 norm over the trainable tensors
  const float* sumsq = nullptr;
  if (a->max_grad_norm > 0.f) {
    FC_CUDA_CHECK(cudaMemsetAsync(w.scalars, 0, sizeof(float), st));
    TRY(fc_sumsq(a->grads, nullptr, a->chunks, a->n_chunks, w.scalars, 1, device, stream));
    sumsq = w.scalars;
  }
  // ---- optimizer.step()
  if (a->optimizer == FC_OPT_ADAMW) {
    TRY(fc_adamw_step(a->params, a->grads, a->opt_state0, a->opt_state1, a->chunks, a->n_chunks, a->lr, a->beta1,
                      a->beta2, a->eps, a->weight_decay, a->step, sumsq, a->max_grad_norm, device, stream));
  } else if (a->optimizer == FC_OPT_SGD) {
    TRY(fc_sgd_step(a->params, a->grads, a->opt_state0, a->chunks, a->n_chunks, a->lr, a->momentum, a->dampening,
                    a->weight_decay, a->nesterov, a->step == 1, sumsq, a->max_grad_norm, device, stream));
  } else if (a->optimizer != FC_OPT_NONE) {
    FC_FAIL(FC_ERR_UNSUPPORTED, "unsupported optimizer id %d", a->optimizer);
  }
  // ---- refresh the bf16 GEMM operands (W + s*A) for the next forward
  if (a->optimizer != FC_OPT_NONE && a->n_prep_layers > 0)
    TRY(fc_prep_weights(a->params, a->operands, a->prep_layers, a->n_prep_layers, a->n_prep_tiles, device, stream));
  return FC_OK;
}
