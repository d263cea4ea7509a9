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

// A lockstep group of clients: up to FC_STEP_MAX_GROUPS instances of the SAME model (one fc_mat_desc), each with its own
// arenas, workspace and inputs, all at the same batch size.  Every tcgen05 GEMM, attention and LayerNorm launch of the
// step covers the whole group (fc_*_grouped); the small kernels (embeddings, heads, losses, optimizer) loop over it.
constexpr int MAXG = FC_STEP_MAX_GROUPS;
static_assert(MAXG <= FC_GEMM_MAX_GROUPS && MAXG <= FC_ATTN_MAX_GROUPS && MAXG <= FC_LN_MAX_GROUPS, "group limits");

struct Ctx {
  const fc_mat_desc* m;
  int G, B, device;
  void* stream;
  const float* params[MAXG];
  float* grads[MAXG];
  const __nv_bfloat16* ops[MAXG];
  const float* droppath[MAXG];    // [2 enc][L][2][B] or null
  Ws w[MAXG];
  const float* p(int g, long long off) const { return off >= 0 ? params[g] + off : nullptr; }
  float* gr(int g, long long off) const { return (off >= 0 && grads[g]) ? grads[g] + off : nullptr; }
  const float* dp(int g, int e, int j, int which) const {
    return droppath[g] ? droppath[g] + (((size_t)e * m->depth + j) * 2 + which) * B : nullptr;
  }
};

// host pointer table over the group, alive until the end of the full expression that builds it
template <typename T> struct Tbl {
  T v[MAXG];
  template <typename F> Tbl(int G, F f) {
    for (int g = 0; g < MAXG; ++g) v[g] = g < G ? f(g) : T();
  }
};
#define TBL(T, expr) (Tbl<T>(c.G, [&](int g) -> T { return (expr); }).v)
#define EACH(g) for (int g = 0; g < c.G; ++g)

int gemm(const Ctx& c, int M, int N, int K, const void* const* A, long long lda, int a_mn, const void* const* Bm,
         long long ldb, int b_mn, int epi, void* const* out, void* const* out2, long long ldo, const float* const* bias,
         const float* const* resid, const float* const* row_scale, int rpg, const void* const* aux,
         const float* const* pos, int patches, int splits, float* const* colsum = nullptr) {
  return fc_gemm_bf16_grouped(c.G, M, N, K, A, lda, a_mn, Bm, ldb, b_mn, epi, out, out2, ldo, bias, resid, row_scale, rpg,
                              aux, pos, patches, 1.0f, splits, colsum, c.device, c.stream);
}

// dW[rows_out, cols_out] += dY[T, rows_out]^T X[T, cols_out]   (both operands MN-major, split-K over tokens)
int gemm_dw(const Ctx& c, int rows_out, int cols_out, int T, const void* const* dY, const void* const* X,
            void* const* dW) {
  return gemm(c, rows_out, cols_out, T, dY, rows_out, 1, X, cols_out, 1, FC_EPI_ATOMIC_F32, dW, nullptr, cols_out,
              nullptr, nullptr, nullptr, 0, nullptr, nullptr, 0, /*splits: auto*/ 0);
}

int encoder_forward(Ctx& c, int e, const float* const* img, const long long* const* ids) {
  const fc_mat_desc* m = c.m;
  const int d = m->d, hid = m->hidden, L = m->depth, H = m->heads, B = c.B;
  const int N = tokens_of(m, e), T = B * N;
#define S(g) c.w[g].enc[e]
  if (e == 0) {
    EACH(g) {
      FC_REQUIRE(img[g] != nullptr, "fc_mat_forward: image input missing");
      TRY(fc_im2col16(img[g], S(g).patches, S(g).x_in[0], c.p(g, m->img_cls), c.p(g, m->img_pos), B, m->in_chans,
                      m->img_size, d, c.device, c.stream));
    }
    TRY(gemm(c, B * m->patches, d, 768, TBL(const void*, S(g).patches), 768, 0, TBL(const void*, c.ops[g] + m->op_pw), 768,
             0, FC_EPI_PATCH, TBL(void*, S(g).x_in[0]), nullptr, d, TBL(const float*, c.p(g, m->img_pb)), nullptr, nullptr,
             0, nullptr, TBL(const float*, c.p(g, m->img_pos)), m->patches, 1));
  } else {
    EACH(g) {
      FC_REQUIRE(ids[g] != nullptr, "fc_mat_forward: token ids missing");
      TRY(fc_text_embed_fwd(ids[g], c.p(g, m->txt_word), c.p(g, m->txt_pos), c.p(g, m->txt_type), c.p(g, m->txt_lnw),
                            c.p(g, m->txt_lnb), 1e-12f, S(g).x_in[0], S(g).mean_e, S(g).rstd_e, B, N, d, c.device,
                            c.stream));
    }
  }
  for (int j = 0; j < L; ++j) {
    const long long* o = m->blk[e][j];
    const long long* op = m->op[e][j];
    TRY(fc_layernorm_fwd_grouped(c.G, TBL(const float*, S(g).x_in[j]), d, TBL(const float*, c.p(g, o[N1W])),
                                 TBL(const float*, c.p(g, o[N1B])), 1e-5f, TBL(void*, S(g).ln1[j]), nullptr,
                                 TBL(float*, S(g).mean1[j]), TBL(float*, S(g).rstd1[j]), T, d, c.device, c.stream));
    TRY(gemm(c, T, 3 * d, d, TBL(const void*, S(g).ln1[j]), d, 0, TBL(const void*, c.ops[g] + op[L_QKV]), d, 0, FC_EPI_BF16,
             TBL(void*, S(g).qkv[j]), nullptr, 3 * d, TBL(const float*, c.p(g, o[QKVB])), nullptr, nullptr, 0, nullptr,
             nullptr, 0, 1));
    TRY(fc_attention_fwd_grouped(c.G, TBL(const void*, S(g).qkv[j]), TBL(void*, S(g).ao[j]), TBL(float*, S(g).lse[j]), B, N,
                                 H, d / H, c.device, c.stream));
    TRY(gemm(c, T, d, d, TBL(const void*, S(g).ao[j]), d, 0, TBL(const void*, c.ops[g] + op[L_PROJ]), d, 0, FC_EPI_RESID,
             TBL(void*, S(g).x_mid[j]), nullptr, d, TBL(const float*, c.p(g, o[PROJB])), TBL(const float*, S(g).x_in[j]),
             TBL(const float*, c.dp(g, e, j, 0)), N, nullptr, nullptr, 0, 1));
    TRY(fc_layernorm_fwd_grouped(c.G, TBL(const float*, S(g).x_mid[j]), d, TBL(const float*, c.p(g, o[N2W])),
                                 TBL(const float*, c.p(g, o[N2B])), 1e-5f, TBL(void*, S(g).ln2[j]), nullptr,
                                 TBL(float*, S(g).mean2[j]), TBL(float*, S(g).rstd2[j]), T, d, c.device, c.stream));
    TRY(gemm(c, T, hid, d, TBL(const void*, S(g).ln2[j]), d, 0, TBL(const void*, c.ops[g] + op[L_FC1]), d, 0, FC_EPI_GELU,
             TBL(void*, S(g).hgrad[j]), TBL(void*, S(g).hact[j]), hid, TBL(const float*, c.p(g, o[FC1B])), nullptr, nullptr,
             0, nullptr, nullptr, 0, 1));
    TRY(gemm(c, T, d, hid, TBL(const void*, S(g).hact[j]), hid, 0, TBL(const void*, c.ops[g] + op[L_FC2]), hid, 0,
             FC_EPI_RESID, TBL(void*, S(g).x_in[j + 1]), nullptr, d, TBL(const float*, c.p(g, o[FC2B])),
             TBL(const float*, S(g).x_mid[j]), TBL(const float*, c.dp(g, e, j, 1)), N, nullptr, nullptr, 0, 1));
  }
  // final norm (eps 1e-6) — only the cls token is consumed by the heads (mome.py:647,658,915)
  TRY(fc_layernorm_fwd_grouped(c.G, TBL(const float*, S(g).x_in[L]), (long long)N * d, TBL(const float*, c.p(g, m->norm_w)),
                               TBL(const float*, c.p(g, m->norm_b)), 1e-6f, nullptr, TBL(float*, S(g).feat),
                               TBL(float*, S(g).mean_f), TBL(float*, S(g).rstd_f), B, d, c.device, c.stream));
  return FC_OK;
}

// dfeat (fp32 [B,d], gradient w.r.t. the final-norm'ed cls token) is in S(g).dfeat
int encoder_backward(Ctx& c, int e, const long long* const* ids) {
  const fc_mat_desc* m = c.m;
  const int d = m->d, hid = m->hidden, L = m->depth, H = m->heads, B = c.B;
  const int N = tokens_of(m, e), T = B * N;
  cudaStream_t st = reinterpret_cast<cudaStream_t>(c.stream);
  EACH(g) {
    FC_CUDA_CHECK(cudaMemsetAsync(S(g).dx, 0, sizeof(float) * (size_t)T * d, st));
    FC_CUDA_CHECK(cudaMemsetAsync(S(g).dxs, 0, sizeof(__nv_bfloat16) * (size_t)T * d, st));
  }
  // final norm backward on the cls rows; dxs = DropPath scale of the last block's mlp branch * dx
  TRY(fc_layernorm_bwd_grouped(c.G, TBL(const void*, S(g).dfeat), 0, d, TBL(const float*, S(g).x_in[L]), (long long)N * d,
                               TBL(const float*, S(g).mean_f), TBL(const float*, S(g).rstd_f),
                               TBL(const float*, c.p(g, m->norm_w)), TBL(float*, S(g).dx), (long long)N * d, 0,
                               TBL(void*, S(g).dxs), (long long)N * d, TBL(const float*, c.dp(g, e, L - 1, 1)), 1,
                               TBL(float*, c.gr(g, m->norm_w)), TBL(float*, c.gr(g, m->norm_b)),
                               TBL(float*, c.gr(g, m->blk[e][L - 1][FC2B])), B, d, c.device, c.stream));
  for (int j = L - 1; j >= 0; --j) {
    const long long* o = m->blk[e][j];
    const long long* op = m->op[e][j];
    // ---- mlp branch:  x_out = x_mid + dp2 * (fc2(gelu(fc1(LN2(x_mid)))))
    // (fc2 bias gradient = column sums of dxs: accumulated by the LayerNorm backward that produced dxs)
    TRY(gemm(c, T, hid, d, TBL(const void*, S(g).dxs), d, 0, TBL(const void*, c.ops[g] + op[L_FC2]), hid, 1, FC_EPI_MULAUX,
             TBL(void*, S(g).d_h), nullptr, hid, nullptr, nullptr, nullptr, 0, TBL(const void*, S(g).hgrad[j]), nullptr, 0, 1,
             TBL(float*, c.gr(g, o[FC1B]))));     // d_h = (dxs W2) * gelu'(pre); fc1 bias grad fused
    TRY(gemm_dw(c, d, hid, T, TBL(const void*, S(g).dxs), TBL(const void*, S(g).hact[j]), TBL(void*, c.gr(g, o[FC2W]))));
    TRY(gemm(c, T, d, hid, TBL(const void*, S(g).d_h), hid, 0, TBL(const void*, c.ops[g] + op[L_FC1]), d, 1, FC_EPI_BF16,
             TBL(void*, S(g).d_ln), nullptr, d, nullptr, nullptr, nullptr, 0, nullptr, nullptr, 0, 1));
    TRY(gemm_dw(c, hid, d, T, TBL(const void*, S(g).d_h), TBL(const void*, S(g).ln2[j]), TBL(void*, c.gr(g, o[FC1W]))));
    TRY(fc_layernorm_bwd_grouped(c.G, TBL(const void*, S(g).d_ln), 1, d, TBL(const float*, S(g).x_mid[j]), d,
                                 TBL(const float*, S(g).mean2[j]), TBL(const float*, S(g).rstd2[j]),
                                 TBL(const float*, c.p(g, o[N2W])), TBL(float*, S(g).dx), d, 1, TBL(void*, S(g).dxs), d,
                                 TBL(const float*, c.dp(g, e, j, 0)), N, TBL(float*, c.gr(g, o[N2W])),
                                 TBL(float*, c.gr(g, o[N2B])), TBL(float*, c.gr(g, o[PROJB])), T, d, c.device, c.stream));
    // ---- attention branch:  x_mid = x_in + dp1 * proj(attn(qkv(LN1(x_in))))
    TRY(gemm(c, T, d, d, TBL(const void*, S(g).dxs), d, 0, TBL(const void*, c.ops[g] + op[L_PROJ]), d, 1, FC_EPI_BF16,
             TBL(void*, S(g).d_ao), nullptr, d, nullptr, nullptr, nullptr, 0, nullptr, nullptr, 0, 1));
    TRY(gemm_dw(c, d, d, T, TBL(const void*, S(g).dxs), TBL(const void*, S(g).ao[j]), TBL(void*, c.gr(g, o[PROJW]))));
    TRY(fc_attention_bwd_grouped(c.G, TBL(const void*, S(g).qkv[j]), TBL(const void*, S(g).ao[j]),
                                 TBL(const void*, S(g).d_ao), TBL(const float*, S(g).lse[j]), TBL(void*, S(g).d_qkv),
                                 TBL(float*, c.gr(g, o[QKVB])), B, N, H, d / H, c.device, c.stream));
    TRY(gemm(c, T, d, 3 * d, TBL(const void*, S(g).d_qkv), 3 * d, 0, TBL(const void*, c.ops[g] + op[L_QKV]), d, 1,
             FC_EPI_BF16, TBL(void*, S(g).d_ln), nullptr, d, nullptr, nullptr, nullptr, 0, nullptr, nullptr, 0, 1));
    TRY(gemm_dw(c, 3 * d, d, T, TBL(const void*, S(g).d_qkv), TBL(const void*, S(g).ln1[j]), TBL(void*, c.gr(g, o[QKVW]))));
    TRY(fc_layernorm_bwd_grouped(c.G, TBL(const void*, S(g).d_ln), 1, d, TBL(const float*, S(g).x_in[j]), d,
                                 TBL(const float*, S(g).mean1[j]), TBL(const float*, S(g).rstd1[j]),
                                 TBL(const float*, c.p(g, o[N1W])), TBL(float*, S(g).dx), d, 1, TBL(void*, S(g).dxs), d,
                                 TBL(const float*, j > 0 ? c.dp(g, e, j - 1, 1) : nullptr), N, TBL(float*, c.gr(g, o[N1W])),
                                 TBL(float*, c.gr(g, o[N1B])),
                                 TBL(float*, j > 0 ? c.gr(g, m->blk[e][j - 1][FC2B]) : nullptr), T, d, c.device, c.stream));
  }
  if (e == 0) {
    EACH(g)
      TRY(fc_patch_bwd_prep(S(g).dx, S(g).dxp, c.gr(g, m->img_pos), c.gr(g, m->img_cls), c.gr(g, m->img_pb), B, m->patches, d,
                            c.device, c.stream));
    TRY(gemm_dw(c, d, 768, B * m->patches, TBL(const void*, S(g).dxp), TBL(const void*, S(g).patches),
                TBL(void*, c.gr(g, m->img_pw))));
  } else {
    EACH(g)
      TRY(fc_text_embed_bwd(S(g).dx, ids[g], c.p(g, m->txt_word), c.p(g, m->txt_pos), c.p(g, m->txt_type),
                            c.p(g, m->txt_lnw), S(g).mean_e, S(g).rstd_e, c.gr(g, m->txt_word), c.gr(g, m->txt_pos),
                            c.gr(g, m->txt_type), c.gr(g, m->txt_lnw), c.gr(g, m->txt_lnb), B, N, d, c.device, c.stream));
  }
#undef S
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

bool is_retrieval(const fc_mat_desc* m, int e, int feat_out);

// =====================================================================================================
// fp32-accurate validation mode (m->precise): the same network with fp32 activations, split-operand GEMMs
// (fc_gemm_split: 3 tcgen05 passes over (hi, lo) bf16 pairs), fp32 FMA attention, exact erff GELU.  One client at a
// time, nothing fused, nothing tuned: it exists to hold the fast path to the reference's fp32 arithmetic at 1e-4.
// =====================================================================================================
struct PEnc {
  float* x_in[FC_MAX_DEPTH + 1];
  float *x_mid[FC_MAX_DEPTH], *ln1[FC_MAX_DEPTH], *ln2[FC_MAX_DEPTH], *qkv[FC_MAX_DEPTH], *ao[FC_MAX_DEPTH],
      *pre[FC_MAX_DEPTH], *act[FC_MAX_DEPTH];
  float *mean1[FC_MAX_DEPTH], *rstd1[FC_MAX_DEPTH], *mean2[FC_MAX_DEPTH], *rstd2[FC_MAX_DEPTH], *lse[FC_MAX_DEPTH];
  float *patches, *mean_e, *rstd_e, *mean_f, *rstd_f, *feat, *featn, *fnorm, *logits;
  float *dx, *d_act, *d_pre, *d_ln, *d_ao, *d_qkv, *dxp, *dfeat;
  __nv_bfloat16* dxp_scratch;
};
struct PWs {
  PEnc enc[2];
  __nv_bfloat16 *sa_hi, *sa_lo, *sb_hi, *sb_lo;      // split scratch of the two GEMM operands
  float *sim, *lse_ws, *da, *db, *dlogits, *scalars, *seg_sumsq;
  size_t bytes;
};

void carve_precise(const fc_mat_desc* m, int B, void* base, PWs* w) {
  Carver c{reinterpret_cast<uint8_t*>(base)};
  const int d = m->d, hid = m->hidden, L = m->depth, H = m->heads;
  int maxC = 1;
  size_t max_op = 0;
  for (int e = 0; e < 2; ++e) {
    if (!m->has_enc[e]) continue;
    PEnc& s = w->enc[e];
    const size_t N = tokens_of(m, e), T = (size_t)B * N;
    const size_t widest = (size_t)(hid > 3 * d ? hid : 3 * d);
    if (T * (widest > 768 ? widest : 768) > max_op) max_op = T * (widest > 768 ? widest : 768);
    for (int j = 0; j <= L; ++j) s.x_in[j] = c.take<float>(T * d);
    for (int j = 0; j < L; ++j) {
      s.x_mid[j] = c.take<float>(T * d);
      s.ln1[j] = c.take<float>(T * d);
      s.ln2[j] = c.take<float>(T * d);
      s.qkv[j] = c.take<float>(T * 3 * d);
      s.ao[j] = c.take<float>(T * d);
      s.pre[j] = c.take<float>(T * hid);
      s.act[j] = c.take<float>(T * hid);
      s.mean1[j] = c.take<float>(T);
      s.rstd1[j] = c.take<float>(T);
      s.mean2[j] = c.take<float>(T);
      s.rstd2[j] = c.take<float>(T);
      s.lse[j] = c.take<float>((size_t)B * H * N);
    }
    s.patches = e == 0 ? c.take<float>((size_t)B * m->patches * 768) : nullptr;
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
    s.d_act = c.take<float>(T * hid);
    s.d_pre = c.take<float>(T * hid);
    s.d_ln = c.take<float>(T * d);
    s.d_ao = c.take<float>(T * d);
    s.d_qkv = c.take<float>(T * 3 * d);
    s.dxp = e == 0 ? c.take<float>((size_t)B * m->patches * d) : nullptr;
    s.dxp_scratch = e == 0 ? c.take<__nv_bfloat16>((size_t)B * m->patches * d) : nullptr;
    s.dfeat = c.take<float>((size_t)B * d);
  }
  w->sa_hi = c.take<__nv_bfloat16>(max_op);
  w->sa_lo = c.take<__nv_bfloat16>(max_op);
  w->sb_hi = c.take<__nv_bfloat16>(max_op);
  w->sb_lo = c.take<__nv_bfloat16>(max_op);
  w->sim = c.take<float>((size_t)B * B);
  w->lse_ws = c.take<float>(2 * (size_t)B);
  w->da = c.take<float>((size_t)B * d);
  w->db = c.take<float>((size_t)B * d);
  w->dlogits = c.take<float>((size_t)B * maxC);
  w->scalars = c.take<float>(8);
  w->seg_sumsq = c.take<float>(FC_MAX_SEGMENTS);
  w->bytes = c.off;
}

struct PCtx {
  const fc_mat_desc* m;
  const float* params;
  float* grads;
  const __nv_bfloat16* ops;      // W_eff hi parts; lo parts at ops + m->op_lo_offset
  int B, device;
  void* stream;
  const float* droppath;
  PWs w;
  const float* p(long long off) const { return off >= 0 ? params + off : nullptr; }
  float* g(long long off) const { return (off >= 0 && grads) ? grads + off : nullptr; }
  const float* dp(int e, int j, int which) const {
    return droppath ? droppath + (((size_t)e * m->depth + j) * 2 + which) * B : nullptr;
  }
};

// out[M, N] = epi( X[M, K] W[N, K]^T )            (forward Linear; W by operand offset)
int plinear(PCtx& c, int M, int N, int K, const float* X, long long w_off, int epi, float* out, const float* bias,
            const float* resid, const float* row_scale, int rpg, const float* pos = nullptr, int patches = 0) {
  TRY(fc_split_bf16(X, c.w.sa_hi, c.w.sa_lo, (long long)M * K, K, nullptr, 0, c.device, c.stream));
  return fc_gemm_split(M, N, K, c.w.sa_hi, c.w.sa_lo, K, 0, c.ops + w_off, c.ops + c.m->op_lo_offset + w_off, K, 0, epi, out, N,
                       bias, resid, row_scale, rpg, pos, patches, 1.0f, 1, c.device, c.stream);
}
// dX[M, K] = dY[M, N] W[N, K]     (dY optionally scaled per row group: the DropPath factor of the branch)
int pdx(PCtx& c, int M, int N, int K, const float* dY, const float* row_scale, int rpg, long long w_off, float* dX) {
  TRY(fc_split_bf16(dY, c.w.sa_hi, c.w.sa_lo, (long long)M * N, N, row_scale, rpg, c.device, c.stream));
  return fc_gemm_split(M, K, N, c.w.sa_hi, c.w.sa_lo, N, 0, c.ops + w_off, c.ops + c.m->op_lo_offset + w_off, K, 1, FC_EPI_F32, dX,
                       K, nullptr, nullptr, nullptr, 0, nullptr, 0, 1.0f, 1, c.device, c.stream);
}
// dW[N, K] += dY[M, N]^T X[M, K]  (sa must already hold the split of (scaled) dY: call right after pdx); db += colsum(dY)
int pdw(PCtx& c, int M, int N, int K, const float* X, float* dW) {
  TRY(fc_split_bf16(X, c.w.sb_hi, c.w.sb_lo, (long long)M * K, K, nullptr, 0, c.device, c.stream));
  return fc_gemm_split(N, K, M, c.w.sa_hi, c.w.sa_lo, N, 1, c.w.sb_hi, c.w.sb_lo, K, 1, FC_EPI_ATOMIC_F32, dW, K, nullptr, nullptr,
                       nullptr, 0, nullptr, 0, 1.0f, 0, c.device, c.stream);
}

int p_encoder_forward(PCtx& c, int e, const float* img, const long long* ids) {
  const fc_mat_desc* m = c.m;
  PEnc& s = c.w.enc[e];
  const int d = m->d, hid = m->hidden, L = m->depth, H = m->heads, B = c.B;
  const int N = tokens_of(m, e), T = B * N;
  if (e == 0) {
    FC_REQUIRE(img != nullptr, "fc_mat_forward: image input missing");
    TRY(fc_im2col16_f32(img, s.patches, B, m->in_chans, m->img_size, c.device, c.stream));
    // cls rows: x[b, 0, :] = cls + pos[0]  (the bf16 patch matrix this call also writes goes to the dW scratch: unused)
    TRY(fc_im2col16(img, s.dxp_scratch ? (void*)c.w.sb_hi : nullptr, s.x_in[0], c.p(m->img_cls), c.p(m->img_pos), B,
                    m->in_chans, m->img_size, d, c.device, c.stream));
    TRY(plinear(c, B * m->patches, d, 768, s.patches, m->op_pw, FC_EPI_PATCH, s.x_in[0], c.p(m->img_pb), nullptr, nullptr, 0,
                c.p(m->img_pos), m->patches));
  } else {
    FC_REQUIRE(ids != nullptr, "fc_mat_forward: token ids missing");
    TRY(fc_text_embed_fwd(ids, c.p(m->txt_word), c.p(m->txt_pos), c.p(m->txt_type), c.p(m->txt_lnw), c.p(m->txt_lnb),
                          1e-12f, s.x_in[0], s.mean_e, s.rstd_e, B, N, d, c.device, c.stream));
  }
  for (int j = 0; j < L; ++j) {
    const long long* o = m->blk[e][j];
    const long long* op = m->op[e][j];
    TRY(fc_layernorm_fwd(s.x_in[j], d, c.p(o[N1W]), c.p(o[N1B]), 1e-5f, nullptr, s.ln1[j], s.mean1[j], s.rstd1[j], T, d,
                         c.device, c.stream));
    TRY(plinear(c, T, 3 * d, d, s.ln1[j], op[L_QKV], FC_EPI_F32, s.qkv[j], c.p(o[QKVB]), nullptr, nullptr, 0));
    TRY(fc_attention_f32_fwd(s.qkv[j], s.ao[j], s.lse[j], B, N, H, c.device, c.stream));
    TRY(plinear(c, T, d, d, s.ao[j], op[L_PROJ], FC_EPI_RESID, s.x_mid[j], c.p(o[PROJB]), s.x_in[j], c.dp(e, j, 0), N));
    TRY(fc_layernorm_fwd(s.x_mid[j], d, c.p(o[N2W]), c.p(o[N2B]), 1e-5f, nullptr, s.ln2[j], s.mean2[j], s.rstd2[j], T, d,
                         c.device, c.stream));
    TRY(plinear(c, T, hid, d, s.ln2[j], op[L_FC1], FC_EPI_F32, s.pre[j], c.p(o[FC1B]), nullptr, nullptr, 0));
    TRY(fc_gelu_f32_fwd(s.pre[j], s.act[j], (long long)T * hid, c.device, c.stream));
    TRY(plinear(c, T, d, hid, s.act[j], op[L_FC2], FC_EPI_RESID, s.x_in[j + 1], c.p(o[FC2B]), s.x_mid[j], c.dp(e, j, 1), N));
  }
  TRY(fc_layernorm_fwd(s.x_in[L], (long long)N * d, c.p(m->norm_w), c.p(m->norm_b), 1e-6f, nullptr, s.feat, s.mean_f,
                       s.rstd_f, B, d, c.device, c.stream));
  return FC_OK;
}

int p_encoder_backward(PCtx& c, int e, const long long* ids) {
  const fc_mat_desc* m = c.m;
  PEnc& s = c.w.enc[e];
  const int d = m->d, hid = m->hidden, L = m->depth, H = m->heads, B = c.B;
  const int N = tokens_of(m, e), T = B * N;
  cudaStream_t st = reinterpret_cast<cudaStream_t>(c.stream);
  FC_CUDA_CHECK(cudaMemsetAsync(s.dx, 0, sizeof(float) * (size_t)T * d, st));
  TRY(fc_layernorm_bwd(s.dfeat, 0, d, s.x_in[L], (long long)N * d, s.mean_f, s.rstd_f, c.p(m->norm_w), s.dx,
                       (long long)N * d, 0, nullptr, 0, nullptr, 1, c.g(m->norm_w), c.g(m->norm_b), nullptr, B, d, c.device,
                       c.stream));
  for (int j = L - 1; j >= 0; --j) {
    const long long* o = m->blk[e][j];
    const long long* op = m->op[e][j];
    // ---- mlp branch: the branch gradient is dp2 * dx (per-sample DropPath factor)
    const float* dp2 = c.dp(e, j, 1);
    TRY(fc_colsum_f32(s.dx, d, T, d, dp2, N, c.g(o[FC2B]), c.device, c.stream));
    TRY(pdx(c, T, d, hid, s.dx, dp2, N, op[L_FC2], s.d_act));                // sa = split(dp2 * dx)
    TRY(pdw(c, T, d, hid, s.act[j], c.g(o[FC2W])));
    TRY(fc_gelu_f32_bwd(s.d_act, s.pre[j], s.d_pre, (long long)T * hid, c.device, c.stream));
    TRY(fc_colsum_f32(s.d_pre, hid, T, hid, nullptr, 0, c.g(o[FC1B]), c.device, c.stream));
    TRY(pdx(c, T, hid, d, s.d_pre, nullptr, 0, op[L_FC1], s.d_ln));
    TRY(pdw(c, T, hid, d, s.ln2[j], c.g(o[FC1W])));
    TRY(fc_layernorm_bwd(s.d_ln, 0, d, s.x_mid[j], d, s.mean2[j], s.rstd2[j], c.p(o[N2W]), s.dx, d, 1, nullptr, 0, nullptr, 1,
                         c.g(o[N2W]), c.g(o[N2B]), nullptr, T, d, c.device, c.stream));
    // ---- attention branch
    const float* dp1 = c.dp(e, j, 0);
    TRY(fc_colsum_f32(s.dx, d, T, d, dp1, N, c.g(o[PROJB]), c.device, c.stream));
    TRY(pdx(c, T, d, d, s.dx, dp1, N, op[L_PROJ], s.d_ao));
    TRY(pdw(c, T, d, d, s.ao[j], c.g(o[PROJW])));
    FC_CUDA_CHECK(cudaMemsetAsync(s.d_qkv, 0, sizeof(float) * (size_t)T * 3 * d, st));
    TRY(fc_attention_f32_bwd(s.qkv[j], s.ao[j], s.d_ao, s.lse[j], s.d_qkv, B, N, H, c.device, c.stream));
    TRY(fc_colsum_f32(s.d_qkv, 3 * d, T, 3 * d, nullptr, 0, c.g(o[QKVB]), c.device, c.stream));
    TRY(pdx(c, T, 3 * d, d, s.d_qkv, nullptr, 0, op[L_QKV], s.d_ln));
    TRY(pdw(c, T, 3 * d, d, s.ln1[j], c.g(o[QKVW])));
    TRY(fc_layernorm_bwd(s.d_ln, 0, d, s.x_in[j], d, s.mean1[j], s.rstd1[j], c.p(o[N1W]), s.dx, d, 1, nullptr, 0, nullptr, 1,
                         c.g(o[N1W]), c.g(o[N1B]), nullptr, T, d, c.device, c.stream));
  }
  if (e == 0) {
    TRY(fc_patch_bwd_prep(s.dx, s.dxp_scratch, c.g(m->img_pos), c.g(m->img_cls), c.g(m->img_pb), B, m->patches, d, c.device,
                          c.stream));
    TRY(fc_drop_cls_rows(s.dx, s.dxp, B, m->patches, d, c.device, c.stream));
    const int M = B * m->patches;
    TRY(fc_split_bf16(s.dxp, c.w.sa_hi, c.w.sa_lo, (long long)M * d, d, nullptr, 0, c.device, c.stream));
    TRY(pdw(c, M, d, 768, s.patches, c.g(m->img_pw)));
  } else {
    TRY(fc_text_embed_bwd(s.dx, ids, c.p(m->txt_word), c.p(m->txt_pos), c.p(m->txt_type), c.p(m->txt_lnw), s.mean_e,
                          s.rstd_e, c.g(m->txt_word), c.g(m->txt_pos), c.g(m->txt_type), c.g(m->txt_lnw), c.g(m->txt_lnb), B,
                          N, d, c.device, c.stream));
  }
  return FC_OK;
}

int p_forward(const fc_mat_desc* m, const float* params, const void* operands, void* workspace, int B, const float* img,
              const long long* ids, const float* droppath, int feat_out, float* out0, float* out1, int device, void* stream) {
  PCtx c{m, params, nullptr, reinterpret_cast<const __nv_bfloat16*>(operands), B, device, stream, droppath, {}};
  carve_precise(m, B, workspace, &c.w);
  float* outs[2] = {out0, out1};
  for (int e = 0; e < 2; ++e) {
    if (!m->has_enc[e]) continue;
    TRY(p_encoder_forward(c, e, img, ids));
    PEnc& s = c.w.enc[e];
    if (is_retrieval(m, e, feat_out)) {
      TRY(fc_l2norm_fwd(s.feat, s.featn, s.fnorm, B, m->d, device, stream));
      if (outs[e])
        FC_CUDA_CHECK(cudaMemcpyAsync(outs[e], s.featn, sizeof(float) * (size_t)B * m->d, cudaMemcpyDeviceToDevice,
                                      reinterpret_cast<cudaStream_t>(stream)));
    } else {
      TRY(fc_head_fwd(s.feat, c.p(m->head_w[e]), c.p(m->head_b[e]), s.logits, B, m->d, m->num_classes[e], device, stream));
      if (outs[e])
        FC_CUDA_CHECK(cudaMemcpyAsync(outs[e], s.logits, sizeof(float) * (size_t)B * m->num_classes[e],
                                      cudaMemcpyDeviceToDevice, reinterpret_cast<cudaStream_t>(stream)));
    }
  }
  return FC_OK;
}

int p_backward(const fc_mat_desc* m, const float* params, const void* operands, void* workspace, int B,
               const long long* ids, const float* droppath, int feat_out, const float* dout0, const float* dout1, float* grads,
               const void* aux_layers, int n_aux_layers, int n_aux_chunks, int device, void* stream) {
  PCtx c{m, params, grads, reinterpret_cast<const __nv_bfloat16*>(operands), B, device, stream, droppath, {}};
  carve_precise(m, B, workspace, &c.w);
  const float* douts[2] = {dout0, dout1};
  for (int e = 0; e < 2; ++e) {
    if (!m->has_enc[e] || douts[e] == nullptr) continue;
    PEnc& s = c.w.enc[e];
    if (is_retrieval(m, e, feat_out)) {
      TRY(fc_l2norm_bwd(douts[e], s.featn, s.fnorm, s.dfeat, B, m->d, device, stream));
    } else {
      TRY(fc_head_bwd(douts[e], s.feat, c.p(m->head_w[e]), c.g(m->head_w[e]), c.g(m->head_b[e]), s.dfeat, B, m->d,
                      m->num_classes[e], device, stream));
    }
    TRY(p_encoder_backward(c, e, ids));
  }
  if (n_aux_layers > 0)
    TRY(fc_aux_grads(params, grads, aux_layers, n_aux_layers, n_aux_chunks, m->aux_trained, device, stream));
  return FC_OK;
}

// forward + loss + backward of one client in the validation mode (the optimizer part is shared with the fast path)
int p_step_fwd_bwd(const fc_mat_desc* m, const fc_step_args* a, int device, void* stream) {
  const int B = a->B, d = m->d;
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  PCtx c{m, a->params, a->grads, reinterpret_cast<const __nv_bfloat16*>(a->operands), B, device, stream, a->droppath, {}};
  carve_precise(m, B, a->workspace, &c.w);
  for (int e = 0; e < 2; ++e)
    if (m->has_enc[e]) TRY(p_encoder_forward(c, e, a->img, a->ids));
  FC_CUDA_CHECK(cudaMemsetAsync(a->grads, 0, sizeof(float) * (size_t)a->arena_floats, st));
  PWs& w = c.w;
  if (a->loss_kind == FC_LOSS_CONTRASTIVE) {
    FC_REQUIRE(m->has_enc[0] && m->has_enc[1], "contrastive loss needs both encoders");
    for (int e = 0; e < 2; ++e) TRY(fc_l2norm_fwd(w.enc[e].feat, w.enc[e].featn, w.enc[e].fnorm, B, d, device, stream));
    TRY(fc_contrastive_loss(w.enc[0].featn, w.enc[1].featn, w.sim, w.lse_ws, w.da, w.db, a->stats, a->stats + 2, B, d,
                            1.0f / 0.07f, 1.0f, device, stream));
    TRY(fc_l2norm_bwd(w.da, w.enc[0].featn, w.enc[0].fnorm, w.enc[0].dfeat, B, d, device, stream));
    TRY(fc_l2norm_bwd(w.db, w.enc[1].featn, w.enc[1].fnorm, w.enc[1].dfeat, B, d, device, stream));
    TRY(p_encoder_backward(c, 0, a->ids));
    TRY(p_encoder_backward(c, 1, a->ids));
  } else {
    const int e = a->loss_kind == FC_LOSS_CE_IMG ? 0 : 1;
    const int C = m->num_classes[e];
    FC_REQUIRE(m->has_enc[e] && C > 0 && a->labels != nullptr, "CE loss needs a classification head and labels");
    PEnc& s = w.enc[e];
    TRY(fc_head_fwd(s.feat, c.p(m->head_w[e]), c.p(m->head_b[e]), s.logits, B, d, C, device, stream));
    TRY(fc_ce_loss(s.logits, a->labels, w.dlogits, a->stats, a->stats + 1, a->stats + 2, B, C, 1.0f, device, stream));
    TRY(fc_head_bwd(w.dlogits, s.feat, c.p(m->head_w[e]), c.g(m->head_w[e]), c.g(m->head_b[e]), s.dfeat, B, d, C, device,
                    stream));
    TRY(p_encoder_backward(c, e, a->ids));
  }
  return FC_OK;
}

void init_ctx(Ctx& c, const fc_mat_desc* m, int G, int B, int device, void* stream) {
  c.m = m; c.G = G; c.B = B; c.device = device; c.stream = stream;
  for (int g = 0; g < MAXG; ++g) { c.params[g] = nullptr; c.grads[g] = nullptr; c.ops[g] = nullptr; c.droppath[g] = nullptr; }
}

}  // namespace

extern "C" long long fc_mat_workspace_bytes(const fc_mat_desc* m, int B) {
  if (check_desc(m, B) != FC_OK) return -1;
  if (m->precise) {
    PWs pw;
    carve_precise(m, B, nullptr, &pw);
    return (long long)pw.bytes;
  }
  Ws w;
  carve(m, B, nullptr, &w);
  return (long long)w.bytes;
}

extern "C" int fc_mat_forward(const fc_mat_desc* m, const float* params, const void* operands, void* workspace, int B,
                              const float* img, const long long* ids, const float* droppath, int feat_out,
                              float* out0, float* out1, int device, void* stream) {
  TRY(check_desc(m, B));
  FcDeviceGuard guard(device);
  if (m->precise)
    return p_forward(m, params, operands, workspace, B, img, ids, droppath, feat_out, out0, out1, device, stream);
  Ctx c;
  init_ctx(c, m, 1, B, device, stream);
  carve(m, B, workspace, &c.w[0]);
  c.params[0] = params; c.ops[0] = reinterpret_cast<const __nv_bfloat16*>(operands); c.droppath[0] = droppath;
  float* outs[2] = {out0, out1};
  for (int e = 0; e < 2; ++e) {
    if (!m->has_enc[e]) continue;
    TRY(encoder_forward(c, e, &img, &ids));
    EncWs& s = c.w[0].enc[e];
    if (is_retrieval(m, e, feat_out)) {
      TRY(fc_l2norm_fwd(s.feat, s.featn, s.fnorm, B, m->d, device, stream));
      if (outs[e])
        FC_CUDA_CHECK(cudaMemcpyAsync(outs[e], s.featn, sizeof(float) * (size_t)B * m->d, cudaMemcpyDeviceToDevice,
                                      reinterpret_cast<cudaStream_t>(stream)));
    } else {
      TRY(fc_head_fwd(s.feat, c.p(0, m->head_w[e]), c.p(0, m->head_b[e]), s.logits, B, m->d, m->num_classes[e], device,
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
  if (m->precise)
    return p_backward(m, params, operands, workspace, B, ids, droppath, feat_out, dout0, dout1, grads, aux_layers,
                      n_aux_layers, n_aux_chunks, device, stream);
  Ctx c;
  init_ctx(c, m, 1, B, device, stream);
  carve(m, B, workspace, &c.w[0]);
  c.params[0] = params; c.grads[0] = grads; c.ops[0] = reinterpret_cast<const __nv_bfloat16*>(operands);
  c.droppath[0] = droppath;
  const float* douts[2] = {dout0, dout1};
  for (int e = 0; e < 2; ++e) {
    if (!m->has_enc[e] || douts[e] == nullptr) continue;
    EncWs& s = c.w[0].enc[e];
    if (is_retrieval(m, e, feat_out)) {
      TRY(fc_l2norm_bwd(douts[e], s.featn, s.fnorm, s.dfeat, B, m->d, device, stream));
    } else {
      TRY(fc_head_bwd(douts[e], s.feat, c.p(0, m->head_w[e]), c.gr(0, m->head_w[e]), c.gr(0, m->head_b[e]), s.dfeat, B,
                      m->d, m->num_classes[e], device, stream));
    }
    TRY(encoder_backward(c, e, &ids));
  }
  if (n_aux_layers > 0)
    TRY(fc_aux_grads(params, grads, aux_layers, n_aux_layers, n_aux_chunks, m->aux_trained, device, stream));
  return FC_OK;
}

// One training step of a lockstep group of clients (ref: one iteration of the batch loop of FedavgClient.update /
// FedproxClient.update, src/client/fedavgclient.py:79-102, for each of the clients the reference's ThreadPoolExecutor
// runs side by side).  Same model, same batch size, same loss kind; everything else is per client.
extern "C" int fc_client_step_group(const fc_mat_desc* m, int n, const fc_step_args* const* args, int device,
                                    void* stream) {
  FC_REQUIRE(n >= 1 && n <= MAXG, "fc_client_step_group: %d clients (1..%d)", n, MAXG);
  FC_REQUIRE(args != nullptr && args[0] != nullptr, "fc_client_step_group: null arguments");
  const int B = args[0]->B, loss_kind = args[0]->loss_kind;
  TRY(check_desc(m, B));
  FC_REQUIRE(loss_kind >= FC_LOSS_CE_IMG && loss_kind <= FC_LOSS_CONTRASTIVE, "bad loss kind %d", loss_kind);
  FcDeviceGuard guard(device);
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  const int d = m->d;
  Ctx c;
  init_ctx(c, m, n, B, device, stream);
  const float* img[MAXG] = {nullptr};
  const long long* ids[MAXG] = {nullptr};
  float* seg_sumsq[MAXG] = {nullptr};
  float* scalars[MAXG] = {nullptr};
  for (int g = 0; g < n; ++g) {
    const fc_step_args* a = args[g];
    FC_REQUIRE(a != nullptr && a->B == B && a->loss_kind == loss_kind, "fc_client_step_group: clients of a group must share "
               "the batch size and the loss kind");
    FC_REQUIRE((a->droppath == nullptr) == (args[0]->droppath == nullptr), "fc_client_step_group: DropPath in all or none");
    FC_REQUIRE(a->n_segments <= FC_MAX_SEGMENTS, "too many parameter segments (%d)", a->n_segments);
    if (m->precise) {
      PWs pw;
      carve_precise(m, B, a->workspace, &pw);
      seg_sumsq[g] = pw.seg_sumsq; scalars[g] = pw.scalars;
    } else {
      carve(m, B, a->workspace, &c.w[g]);
      seg_sumsq[g] = c.w[g].seg_sumsq; scalars[g] = c.w[g].scalars;
    }
    c.params[g] = a->params; c.grads[g] = a->grads; c.ops[g] = reinterpret_cast<const __nv_bfloat16*>(a->operands);
    c.droppath[g] = a->droppath;
    img[g] = a->img; ids[g] = a->ids;
  }

  if (m->precise) {        // validation mode: one client at a time, fp32 activations, split-operand GEMMs
    EACH(g) TRY(p_step_fwd_bwd(m, args[g], device, stream));
  } else {
  // ---- forward
  for (int e = 0; e < 2; ++e)
    if (m->has_enc[e]) TRY(encoder_forward(c, e, img, ids));
  EACH(g)    // optimizer.zero_grad()
    FC_CUDA_CHECK(cudaMemsetAsync(args[g]->grads, 0, sizeof(float) * (size_t)args[g]->arena_floats, st));

  // ---- loss + its gradient w.r.t. the encoder outputs
  if (loss_kind == FC_LOSS_CONTRASTIVE) {
    FC_REQUIRE(m->has_enc[0] && m->has_enc[1], "contrastive loss needs both encoders");
    EACH(g) {
      Ws& w = c.w[g];
      const fc_step_args* a = args[g];
      for (int e = 0; e < 2; ++e) TRY(fc_l2norm_fwd(w.enc[e].feat, w.enc[e].featn, w.enc[e].fnorm, B, d, device, stream));
      TRY(fc_contrastive_loss(w.enc[0].featn, w.enc[1].featn, w.sim, w.lse_ws, w.da, w.db, a->stats, a->stats + 2, B, d,
                              1.0f / 0.07f, 1.0f, device, stream));
      TRY(fc_l2norm_bwd(w.da, w.enc[0].featn, w.enc[0].fnorm, w.enc[0].dfeat, B, d, device, stream));
      TRY(fc_l2norm_bwd(w.db, w.enc[1].featn, w.enc[1].fnorm, w.enc[1].dfeat, B, d, device, stream));
    }
    TRY(encoder_backward(c, 0, ids));
    TRY(encoder_backward(c, 1, ids));
  } else {
    const int e = loss_kind == FC_LOSS_CE_IMG ? 0 : 1;
    const int C = m->num_classes[e];
    EACH(g) {
      const fc_step_args* a = args[g];
      FC_REQUIRE(m->has_enc[e] && C > 0 && a->labels != nullptr, "CE loss needs a classification head and labels");
      EncWs& s = c.w[g].enc[e];
      TRY(fc_head_fwd(s.feat, c.p(g, m->head_w[e]), c.p(g, m->head_b[e]), s.logits, B, d, C, device, stream));
      TRY(fc_ce_loss(s.logits, a->labels, c.w[g].dlogits, a->stats, a->stats + 1, a->stats + 2, B, C, 1.0f, device, stream));
      TRY(fc_head_bwd(c.w[g].dlogits, s.feat, c.p(g, m->head_w[e]), c.gr(g, m->head_w[e]), c.gr(g, m->head_b[e]), s.dfeat,
                      B, d, C, device, stream));
    }
    TRY(encoder_backward(c, e, ids));
  }
  }
  // ---- per client: aux gradients, FedProx term, clipping, optimizer step, operand refresh
  EACH(g) {
    const fc_step_args* a = args[g];
    struct { float* seg_sumsq; float* scalars; } w{seg_sumsq[g], scalars[g]};
    // aux gradients + optimizer step + operand refresh in two launches (fc_opt_fused) whenever nothing needs the whole
    // gradient in memory before the step (no FedProx term, no clipping) — the production configuration
    if (a->n_fused_a + a->n_fused_b > 0 && !m->precise && a->prox_mu <= 0.f && a->max_grad_norm <= 0.f &&
        (a->optimizer == FC_OPT_ADAMW || a->optimizer == FC_OPT_SGD)) {
      TRY(fc_opt_fused(a->params, a->grads, a->opt_state0, a->opt_state1, a->operands, a->fused_a, a->n_fused_a, a->fused_b,
                       a->n_fused_b, a->aux_layers, a->fused_counters, a->optimizer, a->lr, a->beta1, a->beta2, a->eps,
                       a->weight_decay, a->momentum, a->dampening, a->nesterov, a->step, device, stream));
      continue;
    }
    if (a->n_aux_layers > 0)
      TRY(fc_aux_grads(a->params, a->grads, a->aux_layers, a->n_aux_layers, a->n_aux_chunks, m->aux_trained, device, stream));
    // FedProx proximal term (fedproxclient.py:64-67): per-tensor un-squared L2 norms
    if (a->prox_mu > 0.f && a->global_params != nullptr) {
      FC_CUDA_CHECK(cudaMemsetAsync(w.seg_sumsq, 0, sizeof(float) * a->n_segments, st));
      TRY(fc_sumsq(a->params, a->global_params, a->chunks, a->n_chunks, w.seg_sumsq, 0, device, stream));
      TRY(fc_prox_grad(a->grads, a->params, a->global_params, a->chunks, a->n_chunks, w.seg_sumsq, a->n_segments,
                       a->prox_mu, a->stats, a->stats + 2, (float)B, device, stream));
    }
    // clip_grad_norm_ (fedavgclient.py:98-99): global L2 norm over the trainable tensors
    const float* sumsq = nullptr;
    if (a->max_grad_norm > 0.f) {
      FC_CUDA_CHECK(cudaMemsetAsync(w.scalars, 0, sizeof(float), st));
      TRY(fc_sumsq(a->grads, nullptr, a->chunks, a->n_chunks, w.scalars, 1, device, stream));
      sumsq = w.scalars;
    }
    // optimizer.step()
    if (a->optimizer == FC_OPT_ADAMW) {
      TRY(fc_adamw_step(a->params, a->grads, a->opt_state0, a->opt_state1, a->chunks, a->n_chunks, a->lr, a->beta1,
                        a->beta2, a->eps, a->weight_decay, a->step, sumsq, a->max_grad_norm, device, stream));
    } else if (a->optimizer == FC_OPT_SGD) {
      TRY(fc_sgd_step(a->params, a->grads, a->opt_state0, a->chunks, a->n_chunks, a->lr, a->momentum, a->dampening,
                      a->weight_decay, a->nesterov, a->step == 1, sumsq, a->max_grad_norm, device, stream));
    } else if (a->optimizer != FC_OPT_NONE) {
      FC_FAIL(FC_ERR_UNSUPPORTED, "unsupported optimizer id %d", a->optimizer);
    }
    // refresh the bf16 GEMM operands (W + s*A) for the next forward
    if (a->optimizer != FC_OPT_NONE && a->n_prep_layers > 0) {
      if (m->precise)
        TRY(fc_prep_weights_split(a->params, a->operands, reinterpret_cast<__nv_bfloat16*>(a->operands) + m->op_lo_offset,
                                  a->prep_layers, a->n_prep_layers, device, stream));
      else
        TRY(fc_prep_weights(a->params, a->operands, a->prep_layers, a->n_prep_layers, a->n_prep_tiles, device, stream));
    }
  }
  return FC_OK;
}

extern "C" int fc_client_step(const fc_mat_desc* m, const fc_step_args* a, int device, void* stream) {
  FC_REQUIRE(a != nullptr, "fc_client_step: null arguments");
  return fc_client_step_group(m, 1, &a, device, stream);
}
