// Embeddings, heads and losses of the ModalityAgnosticTransformer — the small, HBM/latency-bound ends of
// the client step.
//
//   image side   ImageEmbedding.forward (/root/reference/src/models/mome.py:597-611): the 16x16/16 conv is
//                an im2col GEMM; fc_im2col16 writes the bf16 patch matrix (+ the cls rows), the GEMM's PATCH
//                epilogue adds bias + pos_embed.  Backward: fc_patch_bwd_prep.
//   text side    TextEmbedding.forward (mome.py:632-639) = HF BertEmbeddings: word[ids]+type[0]+pos[l] ->
//                LayerNorm(eps 1e-12): fc_text_embed_fwd / fc_text_embed_bwd.
//   heads        ClassificationHead / RetrievalHead on the final-norm'ed cls token (mome.py:641-659,905-920):
//                fc_head_fwd / fc_head_bwd, fc_l2norm_fwd / fc_l2norm_bwd.
//   losses       nn.CrossEntropyLoss (mean) and ContrastiveLossWithTemperature (tau = 1/0.07, constant because
//                the reference builds a fresh criterion each step; fedavgclient.py:85-95):
//                fc_ce_loss / fc_contrastive_loss, forward + gradient in one launch.
#include "common.cuh"
#include "../../include/fedcola_b200.h"

namespace {

__device__ __forceinline__ uint32_t pack2(float a, float b) {
  __nv_bfloat162 v = __floats2bfloat162_rn(a, b);
  return *reinterpret_cast<uint32_t*>(&v);
}
__device__ __forceinline__ float block_sum(float v, float* s_red) {   // blockDim multiple of 32, <= 1024
  v = warp_sum(v);
  __syncthreads();
  if ((threadIdx.x & 31) == 0) s_red[threadIdx.x >> 5] = v;
  __syncthreads();
  float t = threadIdx.x < (blockDim.x >> 5) ? s_red[threadIdx.x] : 0.f;
  t = warp_sum(t);
  return __shfl_sync(0xffffffffu, t, 0);      // valid in warp 0; broadcast through smem for the others
}
__device__ __forceinline__ float block_sum_all(float v, float* s_red) {
  const float t = block_sum(v, s_red);
  __syncthreads();
  if (threadIdx.x == 0) s_red[32] = t;
  __syncthreads();
  return s_red[32];
}
__device__ __forceinline__ float block_max_all(float v, float* s_red) {
  v = warp_max(v);
  __syncthreads();
  if ((threadIdx.x & 31) == 0) s_red[threadIdx.x >> 5] = v;
  __syncthreads();
  float t = threadIdx.x < (blockDim.x >> 5) ? s_red[threadIdx.x] : -INFINITY;
  t = warp_max(t);
  if (threadIdx.x == 0) s_red[32] = t;
  __syncthreads();
  return s_red[32];
}

// ---- image: im2col for the 16x16 stride-16 conv ---------------------------------------------------------
// img fp32 [B, C, HW, HW] -> patches bf16 [B*P, C*256], K index = c*256 + ph*16 + pw (= conv weight flattening).
// in_chans == 1 is repeated to 3 channels as mome.py:893-894 does.
__global__ void __launch_bounds__(256) im2col16_kernel(const float* __restrict__ img, __nv_bfloat16* __restrict__ patches,
                                                       int B, int Cin, int HW, int grid_p) {
  const int P = grid_p * grid_p;
  const long long total = (long long)B * P * 3 * 16;           // one thread = 16 contiguous pixels of one patch row
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int ph = (int)(i & 15);
    const int c = (int)((i >> 4) % 3);
    const long long row = i / 48;
    const int b = (int)(row / P), t = (int)(row % P);
    const int py = t / grid_p, px = t % grid_p;
    const int cs = Cin == 1 ? 0 : c;
    const float* src = img + (((size_t)b * Cin + cs) * HW + (py * 16 + ph)) * HW + px * 16;
    const float4 a = *reinterpret_cast<const float4*>(src), e = *reinterpret_cast<const float4*>(src + 4);
    const float4 f = *reinterpret_cast<const float4*>(src + 8), h = *reinterpret_cast<const float4*>(src + 12);
    uint4* dst = reinterpret_cast<uint4*>(patches + (size_t)row * 768 + c * 256 + ph * 16);
    dst[0] = make_uint4(pack2(a.x, a.y), pack2(a.z, a.w), pack2(e.x, e.y), pack2(e.z, e.w));
    dst[1] = make_uint4(pack2(f.x, f.y), pack2(f.z, f.w), pack2(h.x, h.y), pack2(h.z, h.w));
  }
}
// x[b, 0, :] = cls_token + pos_embed[0]
__global__ void cls_rows_kernel(float* __restrict__ x, const float* __restrict__ cls, const float* __restrict__ pos,
                                int B, int tokens, int d) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < B * d) {
    const int b = i / d, c = i % d;
    x[(size_t)b * tokens * d + c] = cls[c] + pos[c];
  }
}
// dx fp32 [B, P+1, d] -> dxp bf16 [B*P, d];  dpos[t,:] += sum_b dx[b,t,:];  dcls += sum_b dx[b,0,:];
// dbias += sum_{b, t>=1} dx[b,t,:]
__global__ void __launch_bounds__(128) patch_bwd_prep_kernel(const float* __restrict__ dx, __nv_bfloat16* __restrict__ dxp,
                                                             float* __restrict__ dpos, float* __restrict__ dcls,
                                                             float* __restrict__ dbias, int B, int P, int d) {
  const int t = blockIdx.x;                      // token 0..P
  for (int c = threadIdx.x * 4; c < d; c += blockDim.x * 4) {
    float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
    for (int b = 0; b < B; ++b) {
      const float4 v = *reinterpret_cast<const float4*>(dx + ((size_t)b * (P + 1) + t) * d + c);
      acc.x += v.x; acc.y += v.y; acc.z += v.z; acc.w += v.w;
      if (t > 0)
        *reinterpret_cast<uint2*>(dxp + ((size_t)b * P + (t - 1)) * d + c) = make_uint2(pack2(v.x, v.y), pack2(v.z, v.w));
    }
    float* dp = dpos + (size_t)t * d + c;
    dp[0] += acc.x; dp[1] += acc.y; dp[2] += acc.z; dp[3] += acc.w;      // one block owns row t: no atomics
    float* tgt = t == 0 ? dcls + c : dbias + c;
    atomicAdd(tgt, acc.x); atomicAdd(tgt + 1, acc.y); atomicAdd(tgt + 2, acc.z); atomicAdd(tgt + 3, acc.w);
  }
}

// ---- text embeddings ------------------------------------------------------------------------------------
// one warp per token; e = word[id] + type[0] + pos[l]; x = LN(e; eps)
constexpr int kMaxVec = 8;
__global__ void __launch_bounds__(256) text_embed_fwd_kernel(const long long* __restrict__ ids,
                                                             const float* __restrict__ word, const float* __restrict__ pos,
                                                             const float* __restrict__ type, const float* __restrict__ gamma,
                                                             const float* __restrict__ beta, float eps,
                                                             float* __restrict__ x, float* __restrict__ mean_out,
                                                             float* __restrict__ rstd_out, int rows, int L, int d) {
  const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
  const int nwarps = (gridDim.x * blockDim.x) >> 5, nvec = d >> 2;
  for (int row = warp; row < rows; row += nwarps) {
    const long long id = ids[row];
    const int l = row % L;
    const float4* wr = reinterpret_cast<const float4*>(word + (size_t)id * d);
    const float4* pr = reinterpret_cast<const float4*>(pos + (size_t)l * d);
    const float4* tr = reinterpret_cast<const float4*>(type);
    float4 v[kMaxVec];
    float s = 0.f;
#pragma unroll
    for (int i = 0; i < kMaxVec; ++i) {
      const int c = lane + i * 32;
      if (c < nvec) {
        const float4 a = __ldg(wr + c), b = __ldg(tr + c), e = __ldg(pr + c);
        v[i] = make_float4((a.x + b.x) + e.x, (a.y + b.y) + e.y, (a.z + b.z) + e.z, (a.w + b.w) + e.w);
        s += (v[i].x + v[i].y) + (v[i].z + v[i].w);
      }
    }
    const float mean = warp_sum(s) / d;
    float q = 0.f;
#pragma unroll
    for (int i = 0; i < kMaxVec; ++i) {
      const int c = lane + i * 32;
      if (c < nvec) {
        const float a = v[i].x - mean, b = v[i].y - mean, e = v[i].z - mean, f = v[i].w - mean;
        q += (a * a + b * b) + (e * e + f * f);
      }
    }
    const float rstd = rsqrtf(warp_sum(q) / d + eps);
    if (lane == 0) { mean_out[row] = mean; rstd_out[row] = rstd; }
#pragma unroll
    for (int i = 0; i < kMaxVec; ++i) {
      const int c = lane + i * 32;
      if (c < nvec) {
        const float4 g = __ldg(reinterpret_cast<const float4*>(gamma) + c), b = __ldg(reinterpret_cast<const float4*>(beta) + c);
        reinterpret_cast<float4*>(x + (size_t)row * d)[c] =
            make_float4((v[i].x - mean) * rstd * g.x + b.x, (v[i].y - mean) * rstd * g.y + b.y,
                        (v[i].z - mean) * rstd * g.z + b.z, (v[i].w - mean) * rstd * g.w + b.w);
      }
    }
  }
}
// backward: de = LNbwd(dx); dword[id] += de; dpos[l] += de; dtype[0] += de; dgamma, dbeta
__global__ void __launch_bounds__(256) text_embed_bwd_kernel(const float* __restrict__ dx, const long long* __restrict__ ids,
                                                             const float* __restrict__ word, const float* __restrict__ pos,
                                                             const float* __restrict__ type, const float* __restrict__ gamma,
                                                             const float* __restrict__ mean, const float* __restrict__ rstd,
                                                             float* __restrict__ dword, float* __restrict__ dpos,
                                                             float* __restrict__ dtype, float* __restrict__ dgamma,
                                                             float* __restrict__ dbeta, int rows, int L, int d) {
  extern __shared__ float s_part[];   // [3][wpb][d]
  const int wib = threadIdx.x >> 5, lane = threadIdx.x & 31, wpb = blockDim.x >> 5;
  const int warp = blockIdx.x * wpb + wib, nwarps = gridDim.x * wpb, nvec = d >> 2;
  float4 dg[kMaxVec], db[kMaxVec], dt[kMaxVec];
#pragma unroll
  for (int i = 0; i < kMaxVec; ++i) dg[i] = db[i] = dt[i] = make_float4(0.f, 0.f, 0.f, 0.f);
  for (int row = warp; row < rows; row += nwarps) {
    const long long id = ids[row];
    const int l = row % L;
    const float m = mean[row], r = rstd[row];
    const float4* wr = reinterpret_cast<const float4*>(word + (size_t)id * d);
    const float4* pr = reinterpret_cast<const float4*>(pos + (size_t)l * d);
    const float4* tr = reinterpret_cast<const float4*>(type);
    float4 g[kMaxVec], xh[kMaxVec];
    float s1 = 0.f, s2 = 0.f;
#pragma unroll
    for (int i = 0; i < kMaxVec; ++i) {
      const int c = lane + i * 32;
      if (c < nvec) {
        const float4 a = __ldg(wr + c), b = __ldg(tr + c), e = __ldg(pr + c);
        const float4 dyv = reinterpret_cast<const float4*>(dx + (size_t)row * d)[c];
        const float4 gm = __ldg(reinterpret_cast<const float4*>(gamma) + c);
        xh[i] = make_float4((((a.x + b.x) + e.x) - m) * r, (((a.y + b.y) + e.y) - m) * r,
                            (((a.z + b.z) + e.z) - m) * r, (((a.w + b.w) + e.w) - m) * r);
        g[i] = make_float4(dyv.x * gm.x, dyv.y * gm.y, dyv.z * gm.z, dyv.w * gm.w);
        s1 += (g[i].x + g[i].y) + (g[i].z + g[i].w);
        s2 += (g[i].x * xh[i].x + g[i].y * xh[i].y) + (g[i].z * xh[i].z + g[i].w * xh[i].w);
        dg[i].x += dyv.x * xh[i].x; dg[i].y += dyv.y * xh[i].y; dg[i].z += dyv.z * xh[i].z; dg[i].w += dyv.w * xh[i].w;
        db[i].x += dyv.x; db[i].y += dyv.y; db[i].z += dyv.z; db[i].w += dyv.w;
      }
    }
    const float m1 = warp_sum(s1) / d, m2 = warp_sum(s2) / d;
#pragma unroll
    for (int i = 0; i < kMaxVec; ++i) {
      const int c = lane + i * 32;
      if (c < nvec) {
        const float4 o = make_float4(r * (g[i].x - m1 - xh[i].x * m2), r * (g[i].y - m1 - xh[i].y * m2),
                                     r * (g[i].z - m1 - xh[i].z * m2), r * (g[i].w - m1 - xh[i].w * m2));
        if (id != 0) {   // BertEmbeddings: nn.Embedding(..., padding_idx=pad_token_id=0) -> row 0 gets no gradient
          float* w = dword + (size_t)id * d + c * 4;
          atomicAdd(w, o.x); atomicAdd(w + 1, o.y); atomicAdd(w + 2, o.z); atomicAdd(w + 3, o.w);
        }
        float* p = dpos + (size_t)l * d + c * 4;
        atomicAdd(p, o.x); atomicAdd(p + 1, o.y); atomicAdd(p + 2, o.z); atomicAdd(p + 3, o.w);
        dt[i].x += o.x; dt[i].y += o.y; dt[i].z += o.z; dt[i].w += o.w;
      }
    }
  }
  float* pg = s_part;
  float* pb = s_part + wpb * d;
  float* pt = s_part + 2 * wpb * d;
#pragma unroll
  for (int i = 0; i < kMaxVec; ++i) {
    const int c = lane + i * 32;
    if (c < nvec) {
      reinterpret_cast<float4*>(pg + wib * d)[c] = dg[i];
      reinterpret_cast<float4*>(pb + wib * d)[c] = db[i];
      reinterpret_cast<float4*>(pt + wib * d)[c] = dt[i];
    }
  }
  __syncthreads();
  for (int c = threadIdx.x; c < d; c += blockDim.x) {
    float a = 0.f, b = 0.f, t = 0.f;
    for (int w = 0; w < wpb; ++w) { a += pg[w * d + c]; b += pb[w * d + c]; t += pt[w * d + c]; }
    atomicAdd(dgamma + c, a);
    atomicAdd(dbeta + c, b);
    atomicAdd(dtype + c, t);       // token_type row 0 (row 1 never receives gradient)
  }
}

// ---- classification head: logits = feat W^T + b -----------------------------------------------------------
// one block per sample; feat fp32 [B, d]
__global__ void __launch_bounds__(256) head_fwd_kernel(const float* __restrict__ feat, const float* __restrict__ W,
                                                       const float* __restrict__ bias, float* __restrict__ logits,
                                                       int d, int C) {
  const int b = blockIdx.x, warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nw = blockDim.x >> 5;
  const float* f = feat + (size_t)b * d;
  for (int c = warp; c < C; c += nw) {
    const float* w = W + (size_t)c * d;
    float acc = 0.f;
    for (int k = lane * 4; k < d; k += 128) {
      const float4 a = *reinterpret_cast<const float4*>(f + k), e = __ldg(reinterpret_cast<const float4*>(w + k));
      acc += (a.x * e.x + a.y * e.y) + (a.z * e.z + a.w * e.w);
    }
    acc = warp_sum(acc);
    if (lane == 0) logits[(size_t)b * C + c] = acc + bias[c];
  }
}
// dfeat[b,:] = sum_c dlogits[b,c] W[c,:]   (one block per sample)
__global__ void __launch_bounds__(256) head_bwd_feat_kernel(const float* __restrict__ dlogits, const float* __restrict__ W,
                                                            float* __restrict__ dfeat, int d, int C) {
  const int b = blockIdx.x;
  for (int k = threadIdx.x; k < d; k += blockDim.x) {
    float acc = 0.f;
    for (int c = 0; c < C; ++c) acc += dlogits[(size_t)b * C + c] * __ldg(W + (size_t)c * d + k);
    dfeat[(size_t)b * d + k] = acc;
  }
}
// dW[c,:] += sum_b dlogits[b,c] feat[b,:] ; db[c] += sum_b dlogits[b,c]   (one block per class)
__global__ void __launch_bounds__(256) head_bwd_w_kernel(const float* __restrict__ dlogits, const float* __restrict__ feat,
                                                         float* __restrict__ dW, float* __restrict__ dbias, int B, int d,
                                                         int C) {
  const int c = blockIdx.x;
  for (int k = threadIdx.x; k < d; k += blockDim.x) {
    float acc = 0.f;
    for (int b = 0; b < B; ++b) acc += dlogits[(size_t)b * C + c] * feat[(size_t)b * d + k];
    dW[(size_t)c * d + k] += acc;
  }
  if (threadIdx.x == 0) {
    float acc = 0.f;
    for (int b = 0; b < B; ++b) acc += dlogits[(size_t)b * C + c];
    dbias[c] += acc;
  }
}

// ---- cross entropy (mean) forward + gradient; top-1 correct count ---------------------------------------------
// one block per sample. loss_out[0] += loss_b / B ; dlogits = (softmax - onehot) * gscale / B ; stats[0] += correct
__global__ void __launch_bounds__(128) ce_kernel(const float* __restrict__ logits, const long long* __restrict__ target,
                                                 float* __restrict__ dlogits, float* __restrict__ loss_out,
                                                 float* __restrict__ correct_out, float* __restrict__ loss_sum_out,
                                                 int B, int C, float gscale) {
  __shared__ float s_red[33];
  __shared__ int s_arg;
  const int b = blockIdx.x;
  const float* z = logits + (size_t)b * C;
  float mx = -INFINITY;
  for (int c = threadIdx.x; c < C; c += blockDim.x) mx = fmaxf(mx, z[c]);
  mx = block_max_all(mx, s_red);
  float se = 0.f;
  for (int c = threadIdx.x; c < C; c += blockDim.x) se += expf(z[c] - mx);
  se = block_sum_all(se, s_red);
  const long long t = target[b];
  const float lse = mx + logf(se);
  if (dlogits != nullptr)
    for (int c = threadIdx.x; c < C; c += blockDim.x)
      dlogits[(size_t)b * C + c] = (expf(z[c] - lse) - (c == t ? 1.0f : 0.0f)) * (gscale / B);
  if (threadIdx.x == 0) {
    atomicAdd(loss_out, (lse - z[t]) / B);
    if (loss_sum_out != nullptr) atomicAdd(loss_sum_out, lse - z[t]);
    s_arg = C;
  }
  __syncthreads();
  if (correct_out != nullptr) {     // argmax = first index attaining the max (torch.argmax tie rule)
    for (int c = threadIdx.x; c < C; c += blockDim.x)
      if (z[c] == mx) atomicMin(&s_arg, c);
    __syncthreads();
    if (threadIdx.x == 0 && s_arg == (int)t) atomicAdd(correct_out, 1.0f);
  }
}

// ---- retrieval features: out = v / ||v|| ------------------------------------------------------------------
__global__ void __launch_bounds__(128) l2norm_fwd_kernel(const float* __restrict__ v, float* __restrict__ out,
                                                         float* __restrict__ norm_out, int d) {
  __shared__ float s_red[33];
  const int b = blockIdx.x;
  float acc = 0.f;
  for (int k = threadIdx.x; k < d; k += blockDim.x) { const float x = v[(size_t)b * d + k]; acc += x * x; }
  const float nrm = sqrtf(block_sum_all(acc, s_red));
  for (int k = threadIdx.x; k < d; k += blockDim.x) out[(size_t)b * d + k] = v[(size_t)b * d + k] / nrm;
  if (threadIdx.x == 0) norm_out[b] = nrm;
}
// dv = (dout - out * <dout, out>) / ||v||
__global__ void __launch_bounds__(128) l2norm_bwd_kernel(const float* __restrict__ dout, const float* __restrict__ out,
                                                         const float* __restrict__ norm, float* __restrict__ dv, int d) {
  __shared__ float s_red[33];
  const int b = blockIdx.x;
  float acc = 0.f;
  for (int k = threadIdx.x; k < d; k += blockDim.x) acc += dout[(size_t)b * d + k] * out[(size_t)b * d + k];
  const float dot = block_sum_all(acc, s_red);
  const float inv = 1.0f / norm[b];
  for (int k = threadIdx.x; k < d; k += blockDim.x)
    dv[(size_t)b * d + k] = (dout[(size_t)b * d + k] - out[(size_t)b * d + k] * dot) * inv;
}

// ---- contrastive loss with temperature: 0.5*(CE(tau a b^T) + CE(tau b a^T)), labels = arange(B) ---------------------
// pass 1: S[i,j] = tau <a_i, b_j>   (B <= 1024; one block per row)
__global__ void __launch_bounds__(256) sim_kernel(const float* __restrict__ a, const float* __restrict__ bm,
                                                  float* __restrict__ S, int B, int d, float tau) {
  const int i = blockIdx.x, warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nw = blockDim.x >> 5;
  for (int j = warp; j < B; j += nw) {
    float acc = 0.f;
    for (int k = lane * 4; k < d; k += 128) {
      const float4 x = *reinterpret_cast<const float4*>(a + (size_t)i * d + k);
      const float4 y = *reinterpret_cast<const float4*>(bm + (size_t)j * d + k);
      acc += (x.x * y.x + x.y * y.y) + (x.z * y.z + x.w * y.w);
    }
    acc = warp_sum(acc);
    if (lane == 0) S[(size_t)i * B + j] = acc * tau;
  }
}
// pass 2: row / column log-sum-exp of S; loss += 0.5/B * ((lse_row_i - S_ii) + (lse_col_i - S_ii))
__global__ void __launch_bounds__(128) lse_kernel(const float* __restrict__ S, float* __restrict__ lse_row,
                                                  float* __restrict__ lse_col, float* __restrict__ loss_out,
                                                  float* __restrict__ loss_sum_out, int B) {
  __shared__ float s_red[33];
  const int i = blockIdx.x;
  float mr = -INFINITY, mc = -INFINITY;
  for (int j = threadIdx.x; j < B; j += blockDim.x) {
    mr = fmaxf(mr, S[(size_t)i * B + j]);
    mc = fmaxf(mc, S[(size_t)j * B + i]);
  }
  mr = block_max_all(mr, s_red);
  mc = block_max_all(mc, s_red);
  float sr = 0.f, sc = 0.f;
  for (int j = threadIdx.x; j < B; j += blockDim.x) {
    sr += expf(S[(size_t)i * B + j] - mr);
    sc += expf(S[(size_t)j * B + i] - mc);
  }
  sr = block_sum_all(sr, s_red);
  sc = block_sum_all(sc, s_red);
  if (threadIdx.x == 0) {
    const float lr = mr + logf(sr), lc = mc + logf(sc), sii = S[(size_t)i * B + i];
    lse_row[i] = lr;
    lse_col[i] = lc;
    atomicAdd(loss_out, 0.5f * ((lr - sii) + (lc - sii)) / B);
    if (loss_sum_out != nullptr) atomicAdd(loss_sum_out, 0.5f * ((lr - sii) + (lc - sii)));
  }
}
// pass 3: G[i,j] = dL/dS[i,j] * tau = tau * gscale * 0.5/B * (softmax_row + softmax_col - 2*delta_ij);
//         da_i = sum_j G[i,j] b_j ; db_j = sum_i G[i,j] a_i          (one block per row i of da, and of db)
__global__ void __launch_bounds__(128) contrastive_grad_kernel(const float* __restrict__ S, const float* __restrict__ lse_row,
                                                               const float* __restrict__ lse_col,
                                                               const float* __restrict__ a, const float* __restrict__ bm,
                                                               float* __restrict__ da, float* __restrict__ db, int B, int d,
                                                               float coef) {
  extern __shared__ float s_g[];    // [2][B]: weights for da_i (row i of G) and for db_i (column i of G)
  const int i = blockIdx.x;
  for (int j = threadIdx.x; j < B; j += blockDim.x) {
    const float sij = S[(size_t)i * B + j], sji = S[(size_t)j * B + i];
    const float dlt = (i == j) ? 2.0f : 0.0f;
    s_g[j] = coef * (expf(sij - lse_row[i]) + expf(sij - lse_col[j]) - dlt);          // G[i,j]
    s_g[B + j] = coef * (expf(sji - lse_row[j]) + expf(sji - lse_col[i]) - dlt);      // G[j,i]
  }
  __syncthreads();
  for (int k = threadIdx.x; k < d; k += blockDim.x) {
    float x = 0.f, y = 0.f;
    for (int j = 0; j < B; ++j) {
      x += s_g[j] * bm[(size_t)j * d + k];
      y += s_g[B + j] * a[(size_t)j * d + k];
    }
    da[(size_t)i * d + k] = x;
    db[(size_t)i * d + k] = y;
  }
}

}  // namespace

extern "C" int fc_im2col16(const float* img, void* patches_bf16, float* x, const float* cls_token,
                           const float* pos_embed, int B, int in_chans, int img_size, int d, int device,
                           void* stream) {
  FC_REQUIRE(img_size % 16 == 0 && (in_chans == 3 || in_chans == 1), "fc_im2col16: img_size %% 16, in_chans in {1,3}");
  if (B <= 0) return FC_OK;
  FcDeviceGuard guard(device);
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  const int gp = img_size / 16;
  const long long total = (long long)B * gp * gp * 48;
  int grid = (int)((total + 255) / 256);
  const int cap = fc_num_sms(device) * 16;
  if (grid > cap) grid = cap;
  im2col16_kernel<<<grid, 256, 0, st>>>(img, reinterpret_cast<__nv_bfloat16*>(patches_bf16), B, in_chans, img_size, gp);
  FC_LAUNCH_CHECK();
  if (x != nullptr) {
    cls_rows_kernel<<<(B * d + 255) / 256, 256, 0, st>>>(x, cls_token, pos_embed, B, gp * gp + 1, d);
    FC_LAUNCH_CHECK();
  }
  return FC_OK;
}

extern "C" int fc_patch_bwd_prep(const float* dx, void* dxp_bf16, float* dpos, float* dcls, float* dbias, int B,
                                 int patches, int d, int device, void* stream) {
  FC_REQUIRE(d % 4 == 0, "fc_patch_bwd_prep: d %% 4");
  if (B <= 0) return FC_OK;
  FcDeviceGuard guard(device);
  patch_bwd_prep_kernel<<<patches + 1, 128, 0, reinterpret_cast<cudaStream_t>(stream)>>>(
      dx, reinterpret_cast<__nv_bfloat16*>(dxp_bf16), dpos, dcls, dbias, B, patches, d);
  FC_LAUNCH_CHECK();
  return FC_OK;
}

extern "C" int fc_text_embed_fwd(const long long* ids, const float* word, const float* pos, const float* type,
                                 const float* gamma, const float* beta, float eps, float* x, float* mean, float* rstd,
                                 int B, int L, int d, int device, void* stream) {
  FC_REQUIRE(d % 4 == 0 && d <= kMaxVec * 128, "fc_text_embed_fwd: d=%d unsupported", d);
  const int rows = B * L;
  if (rows <= 0) return FC_OK;
  FcDeviceGuard guard(device);
  int grid = (rows + 7) / 8;
  const int cap = fc_num_sms(device) * 8;
  if (grid > cap) grid = cap;
  text_embed_fwd_kernel<<<grid, 256, 0, reinterpret_cast<cudaStream_t>(stream)>>>(ids, word, pos, type, gamma, beta, eps,
                                                                                x, mean, rstd, rows, L, d);
  FC_LAUNCH_CHECK();
  return FC_OK;
}

extern "C" int fc_text_embed_bwd(const float* dx, const long long* ids, const float* word, const float* pos,
                                 const float* type, const float* gamma, const float* mean, const float* rstd,
                                 float* dword, float* dpos, float* dtype, float* dgamma, float* dbeta, int B, int L,
                                 int d, int device, void* stream) {
  FC_REQUIRE(d % 4 == 0 && d <= kMaxVec * 128, "fc_text_embed_bwd: d=%d unsupported", d);
  const int rows = B * L;
  if (rows <= 0) return FC_OK;
  FcDeviceGuard guard(device);
  int grid = (rows + 7) / 8;
  const int cap = fc_num_sms(device) * 2;
  if (grid > cap) grid = cap;
  const size_t smem = sizeof(float) * 3 * 8 * d;
  if (smem > 48 * 1024)
    FC_CUDA_CHECK(cudaFuncSetAttribute(text_embed_bwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  text_embed_bwd_kernel<<<grid, 256, smem, reinterpret_cast<cudaStream_t>(stream)>>>(
      dx, ids, word, pos, type, gamma, mean, rstd, dword, dpos, dtype, dgamma, dbeta, rows, L, d);
  FC_LAUNCH_CHECK();
  return FC_OK;
}

extern "C" int fc_head_fwd(const float* feat, const float* W, const float* bias, float* logits, int B, int d, int C,
                           int device, void* stream) {
  FC_REQUIRE(d % 4 == 0, "fc_head_fwd: d %% 4");
  if (B <= 0) return FC_OK;
  FcDeviceGuard guard(device);
  head_fwd_kernel<<<B, 256, 0, reinterpret_cast<cudaStream_t>(stream)>>>(feat, W, bias, logits, d, C);
  FC_LAUNCH_CHECK();
  return FC_OK;
}

extern "C" int fc_head_bwd(const float* dlogits, const float* feat, const float* W, float* dW, float* dbias,
                           float* dfeat, int B, int d, int C, int device, void* stream) {
  if (B <= 0) return FC_OK;
  FcDeviceGuard guard(device);
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  head_bwd_feat_kernel<<<B, 256, 0, st>>>(dlogits, W, dfeat, d, C);
  FC_LAUNCH_CHECK();
  head_bwd_w_kernel<<<C, 256, 0, st>>>(dlogits, feat, dW, dbias, B, d, C);
  FC_LAUNCH_CHECK();
  return FC_OK;
}

extern "C" int fc_ce_loss(const float* logits, const long long* target, float* dlogits, float* loss_out,
                          float* correct_out, float* loss_sum_out, int B, int C, float grad_scale, int device,
                          void* stream) {
  if (B <= 0) return FC_OK;
  FcDeviceGuard guard(device);
  ce_kernel<<<B, 128, 0, reinterpret_cast<cudaStream_t>(stream)>>>(logits, target, dlogits, loss_out, correct_out,
                                                                 loss_sum_out, B, C, grad_scale);
  FC_LAUNCH_CHECK();
  return FC_OK;
}

extern "C" int fc_l2norm_fwd(const float* v, float* out, float* norm, int B, int d, int device, void* stream) {
  if (B <= 0) return FC_OK;
  FcDeviceGuard guard(device);
  l2norm_fwd_kernel<<<B, 128, 0, reinterpret_cast<cudaStream_t>(stream)>>>(v, out, norm, d);
  FC_LAUNCH_CHECK();
  return FC_OK;
}

extern "C" int fc_l2norm_bwd(const float* dout, const float* out, const float* norm, float* dv, int B, int d,
                             int device, void* stream) {
  if (B <= 0) return FC_OK;
  FcDeviceGuard guard(device);
  l2norm_bwd_kernel<<<B, 128, 0, reinterpret_cast<cudaStream_t>(stream)>>>(dout, out, norm, dv, d);
  FC_LAUNCH_CHECK();
  return FC_OK;
}

extern "C" int fc_contrastive_loss(const float* a, const float* b, float* sim_ws, float* lse_ws, float* da, float* db,
                                   float* loss_out, float* loss_sum_out, int B, int d, float tau, float grad_scale,
                                   int device, void* stream) {
  FC_REQUIRE(d % 4 == 0 && B <= 4096, "fc_contrastive_loss: d %% 4, B <= 4096");
  if (B <= 0) return FC_OK;
  FcDeviceGuard guard(device);
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  sim_kernel<<<B, 256, 0, st>>>(a, b, sim_ws, B, d, tau);
  FC_LAUNCH_CHECK();
  lse_kernel<<<B, 128, 0, st>>>(sim_ws, lse_ws, lse_ws + B, loss_out, loss_sum_out, B);
  FC_LAUNCH_CHECK();
  if (da != nullptr) {
    contrastive_grad_kernel<<<B, 128, sizeof(float) * 2 * B, st>>>(sim_ws, lse_ws, lse_ws + B, a, b, da, db, B, d,
                                                                  tau * grad_scale * 0.5f / B);
    FC_LAUNCH_CHECK();
  }
  return FC_OK;
}
