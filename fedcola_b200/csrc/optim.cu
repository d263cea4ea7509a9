// Client-side optimizer step and the elementwise/reduction kernels around it — all HBM-bound, vectorised
// (128-bit) and table-driven so that ONE launch covers the client's whole flat arena.
//
// Replaces, in FedavgClient.update / FedproxClient.update
// (/root/reference/src/client/fedavgclient.py:63,97-100; fedproxclient.py:64-67):
//   torch.optim.AdamW / SGD .step()          -> fc_adamw_step / fc_sgd_step   (28 B / 12-16 B per param)
//   torch.nn.utils.clip_grad_norm_           -> fc_sumsq + the grad_scale read by the step kernels
//   FedProx  mu*0.5*sum_i ||p_i - g_i||_2    -> fc_prox_sumsq + fc_prox_grad  (per-tensor norms)
// and produces what the next forward needs:
//   fc_prep_weights : bf16 W_eff = W + s*A and its transpose for the tcgen05 GEMMs (the aux mix of
//                     CrossModalReparamLinear.forward, src/models/mome.py:58-60, never runs as a separate
//                     fp32 elementwise op)
//   fc_aux_grads    : dA = s*dW_eff, ds = <dW_eff, A>   (autograd of mome.py:59)
//   fc_colsum_bf16  : bias gradients (column sums of the bf16 activation gradients)
#include "common.cuh"
#include "../../include/fedcola_b200.h"

namespace {

constexpr int kChunk = 2048;   // floats per (segment, chunk) work item = 256 threads x 2 float4

// A chunk table entry: [offset (floats, multiple of 4), length (<= kChunk), segment id]
struct Chunk {
  long long off;
  int len;
  int seg;
};

__device__ __forceinline__ float4 ld4(const float* p) { return *reinterpret_cast<const float4*>(p); }
__device__ __forceinline__ void st4(float* p, const float4& v) { *reinterpret_cast<float4*>(p) = v; }

// ---- AdamW ------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) adamw_kernel(float* __restrict__ p, const float* __restrict__ g,
                                                    float* __restrict__ m, float* __restrict__ v,
                                                    const Chunk* __restrict__ chunks, int n_chunks, float lr,
                                                    float beta1, float beta2, float eps, float wd, float bc1,
                                                    float bc2_sqrt, const float* __restrict__ grad_sumsq,
                                                    float max_norm) {
  float gs = 1.0f;
  if (grad_sumsq != nullptr) {        // clip_grad_norm_: coef = clamp(max_norm / (total_norm + 1e-6), max=1)
    const float coef = max_norm / (sqrtf(__ldg(grad_sumsq)) + 1e-6f);
    gs = fminf(coef, 1.0f);
  }
  const float step_size = lr / bc1;
  for (int c = blockIdx.x; c < n_chunks; c += gridDim.x) {
    const Chunk ch = chunks[c];
    for (int i = threadIdx.x * 4; i < ch.len; i += blockDim.x * 4) {
      const long long o = ch.off + i;
      float4 pv = ld4(p + o), gv = ld4(g + o), mv = ld4(m + o), vv = ld4(v + o);
      float* pp = &pv.x; float* gp = &gv.x; float* mp = &mv.x; float* vp = &vv.x;
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        const float grad = gp[k] * gs;
        float w = pp[k] * (1.0f - lr * wd);                      // param.mul_(1 - lr*wd)
        mp[k] = mp[k] + (grad - mp[k]) * (1.0f - beta1);          // exp_avg.lerp_(grad, 1-beta1)
        vp[k] = vp[k] * beta2 + (1.0f - beta2) * grad * grad;     // exp_avg_sq.mul_(b2).addcmul_(g,g,1-b2)
        const float denom = sqrtf(vp[k]) / bc2_sqrt + eps;
        pp[k] = w - step_size * (mp[k] / denom);                  // param.addcdiv_(exp_avg, denom, -step_size)
      }
      st4(p + o, pv); st4(m + o, mv); st4(v + o, vv);
    }
  }
}

// ---- SGD (momentum / nesterov / weight decay as torch.optim.SGD) ------------------------------------------
__global__ void __launch_bounds__(256) sgd_kernel(float* __restrict__ p, const float* __restrict__ g,
                                                  float* __restrict__ buf, const Chunk* __restrict__ chunks,
                                                  int n_chunks, float lr, float momentum, float dampening, float wd,
                                                  int nesterov, int first_step, const float* __restrict__ grad_sumsq,
                                                  float max_norm) {
  float gs = 1.0f;
  if (grad_sumsq != nullptr) gs = fminf(max_norm / (sqrtf(__ldg(grad_sumsq)) + 1e-6f), 1.0f);
  for (int c = blockIdx.x; c < n_chunks; c += gridDim.x) {
    const Chunk ch = chunks[c];
    for (int i = threadIdx.x * 4; i < ch.len; i += blockDim.x * 4) {
      const long long o = ch.off + i;
      float4 pv = ld4(p + o), gv = ld4(g + o), bv = make_float4(0.f, 0.f, 0.f, 0.f);
      if (momentum != 0.0f && !first_step) bv = ld4(buf + o);
      float* pp = &pv.x; float* gp = &gv.x; float* bp = &bv.x;
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        float grad = gp[k] * gs;
        if (wd != 0.0f) grad += wd * pp[k];
        if (momentum != 0.0f) {
          bp[k] = first_step ? grad : momentum * bp[k] + (1.0f - dampening) * grad;
          grad = nesterov ? grad + momentum * bp[k] : bp[k];
        }
        pp[k] -= lr * grad;
      }
      st4(p + o, pv);
      if (momentum != 0.0f) st4(buf + o, bv);
    }
  }
}

// ---- sum of squares of a (or a-b) per segment -----------------------------------------------------------
// out[seg] += sum (a-b)^2 over the chunk (b may be null).  Hierarchical: thread -> warp shuffle -> block smem
// -> one atomicAdd per chunk.
__global__ void __launch_bounds__(256) sumsq_kernel(const float* __restrict__ a, const float* __restrict__ b,
                                                    const Chunk* __restrict__ chunks, int n_chunks,
                                                    float* __restrict__ out, int single_output) {
  __shared__ float s_red[8];
  for (int c = blockIdx.x; c < n_chunks; c += gridDim.x) {
    const Chunk ch = chunks[c];
    float acc = 0.f;
    for (int i = threadIdx.x * 4; i < ch.len; i += blockDim.x * 4) {
      float4 x = ld4(a + ch.off + i);
      if (b != nullptr) {
        const float4 y = ld4(b + ch.off + i);
        x.x -= y.x; x.y -= y.y; x.z -= y.z; x.w -= y.w;
      }
      acc += (x.x * x.x + x.y * x.y) + (x.z * x.z + x.w * x.w);
    }
    acc = warp_sum(acc);
    if ((threadIdx.x & 31) == 0) s_red[threadIdx.x >> 5] = acc;
    __syncthreads();
    if (threadIdx.x < 32) {
      float t = threadIdx.x < 8 ? s_red[threadIdx.x] : 0.f;
      t = warp_sum(t);
      if (threadIdx.x == 0) atomicAdd(out + (single_output ? 0 : ch.seg), t);
    }
    __syncthreads();
  }
}

// FedProx: grad += mu*0.5*(p - pg)/||p - pg||_seg  (zero where the norm is zero, as torch's norm backward)
__global__ void __launch_bounds__(256) prox_grad_kernel(float* __restrict__ grad, const float* __restrict__ p,
                                                        const float* __restrict__ pg,
                                                        const Chunk* __restrict__ chunks, int n_chunks,
                                                        const float* __restrict__ seg_sumsq, float mu) {
  for (int c = blockIdx.x; c < n_chunks; c += gridDim.x) {
    const Chunk ch = chunks[c];
    const float nrm = sqrtf(__ldg(seg_sumsq + ch.seg));
    if (nrm == 0.0f) continue;
    const float k = 0.5f * mu / nrm;
    for (int i = threadIdx.x * 4; i < ch.len; i += blockDim.x * 4) {
      const long long o = ch.off + i;
      float4 gv = ld4(grad + o);
      const float4 a = ld4(p + o), b = ld4(pg + o);
      gv.x += k * (a.x - b.x); gv.y += k * (a.y - b.y); gv.z += k * (a.z - b.z); gv.w += k * (a.w - b.w);
      st4(grad + o, gv);
    }
  }
}

// prox loss value: out[0] += mu*0.5*sum_seg sqrt(sumsq[seg])
__global__ void prox_loss_kernel(const float* __restrict__ seg_sumsq, int n_seg, float mu, float* __restrict__ out,
                                 float* __restrict__ out_weighted, float weight) {
  float acc = 0.f;
  for (int i = threadIdx.x; i < n_seg; i += blockDim.x) acc += sqrtf(seg_sumsq[i]);
  acc = warp_sum(acc);
  __shared__ float s[32];
  if ((threadIdx.x & 31) == 0) s[threadIdx.x >> 5] = acc;
  __syncthreads();
  if (threadIdx.x < 32) {
    float t = threadIdx.x < (blockDim.x >> 5) ? s[threadIdx.x] : 0.f;
    t = warp_sum(t);
    if (threadIdx.x == 0) {
      atomicAdd(out, 0.5f * mu * t);
      if (out_weighted != nullptr) atomicAdd(out_weighted, 0.5f * mu * t * weight);
    }
  }
}

// ---- bf16 operand preparation: W_eff = W + s*A  ->  bf16 [N,K] and bf16 [K,N] ----------------------------
struct PrepLayer {
  long long w_off, a_off, s_off;     // float offsets into the param arena (a_off/s_off = -1: plain Linear)
  long long dst_off, dstT_off;       // bf16 element offsets into the operand arena (dstT_off = -1: skip)
  int rows, cols;                    // W is [rows=N, cols=K]
  int tile_start;                    // prefix over 32x32 tiles
};

__global__ void __launch_bounds__(256) prep_weights_kernel(const float* __restrict__ params,
                                                           __nv_bfloat16* __restrict__ wb,
                                                           const PrepLayer* __restrict__ layers, int n_layers,
                                                           int n_tiles) {
  __shared__ float tile[32][33];
  __shared__ int s_layer;
  for (int tix = blockIdx.x; tix < n_tiles; tix += gridDim.x) {
    if (threadIdx.x == 0) {
      int lo = 0, hi = n_layers;
      while (hi - lo > 1) {
        const int mid = (lo + hi) >> 1;
        if (layers[mid].tile_start <= tix) lo = mid; else hi = mid;
      }
      s_layer = lo;
    }
    __syncthreads();
    const PrepLayer L = layers[s_layer];
    const int tiles_x = (L.cols + 31) >> 5;
    const int lt = tix - L.tile_start;
    const int r0 = (lt / tiles_x) * 32, c0 = (lt % tiles_x) * 32;
    const float s = L.a_off >= 0 ? __ldg(params + L.s_off) : 0.f;
    const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;     // 32 x 8
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      const int r = r0 + ty + k * 8, c = c0 + tx;
      float w = 0.f;
      if (r < L.rows && c < L.cols) {
        const size_t o = (size_t)r * L.cols + c;
        w = params[L.w_off + o];
        if (L.a_off >= 0) w = w + s * params[L.a_off + o];        // weight + cross_modal_scale * aux_weight
        wb[L.dst_off + o] = __float2bfloat16_rn(w);
      }
      tile[ty + k * 8][tx] = w;
    }
    __syncthreads();
    if (L.dstT_off >= 0) {
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        const int c = c0 + ty + k * 8, r = r0 + tx;               // transposed: row index = original column
        if (r < L.rows && c < L.cols) wb[L.dstT_off + (size_t)c * L.rows + r] = __float2bfloat16_rn(tile[tx][ty + k * 8]);
      }
    }
    __syncthreads();
  }
}

// ---- aux gradients: dA = s * dW_eff ; ds += <dW_eff, A> ---------------------------------------------------
struct AuxLayer {
  long long w_off, a_off, s_off;   // offsets valid for both the param and the grad arena (same layout)
  long long numel;
  int chunk_start;                 // prefix over kChunk-sized chunks
};

__global__ void __launch_bounds__(256) aux_grads_kernel(const float* __restrict__ params, float* __restrict__ grads,
                                                        const AuxLayer* __restrict__ layers, int n_layers,
                                                        int n_chunks, int aux_trained) {
  __shared__ float s_red[8];
  __shared__ int s_layer;
  for (int cix = blockIdx.x; cix < n_chunks; cix += gridDim.x) {
    if (threadIdx.x == 0) {
      int lo = 0, hi = n_layers;
      while (hi - lo > 1) {
        const int mid = (lo + hi) >> 1;
        if (layers[mid].chunk_start <= cix) lo = mid; else hi = mid;
      }
      s_layer = lo;
    }
    __syncthreads();
    const AuxLayer L = layers[s_layer];
    const long long base = (long long)(cix - L.chunk_start) * kChunk;
    const float s = __ldg(params + L.s_off);
    float acc = 0.f;
    for (int i = threadIdx.x * 4; i < kChunk; i += blockDim.x * 4) {
      const long long o = base + i;
      if (o < L.numel) {               // numel is a multiple of 4 (rows*cols of Linear layers)
        const float4 dw = ld4(grads + L.w_off + o);
        const float4 a = ld4(params + L.a_off + o);
        acc += (dw.x * a.x + dw.y * a.y) + (dw.z * a.z + dw.w * a.w);
        if (aux_trained) st4(grads + L.a_off + o, make_float4(s * dw.x, s * dw.y, s * dw.z, s * dw.w));
      }
    }
    acc = warp_sum(acc);
    if ((threadIdx.x & 31) == 0) s_red[threadIdx.x >> 5] = acc;
    __syncthreads();
    if (threadIdx.x < 32) {
      float t = threadIdx.x < 8 ? s_red[threadIdx.x] : 0.f;
      t = warp_sum(t);
      if (threadIdx.x == 0) atomicAdd(grads + L.s_off, t);
    }
    __syncthreads();
  }
}

// ---- column sums of a bf16 matrix [rows, n] -> out[n] += ... (bias gradients) ---------------------------------
__global__ void __launch_bounds__(256) colsum_bf16_kernel(const __nv_bfloat16* __restrict__ x, long long ld, int rows,
                                                          int n, int rows_per_block, float* __restrict__ out) {
  __shared__ float s_part[8][256 + 8];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int c0 = blockIdx.x * 256 + lane * 8;
  const int r_begin = blockIdx.y * rows_per_block, r_end = min(rows, r_begin + rows_per_block);
  float acc[8];
#pragma unroll
  for (int k = 0; k < 8; ++k) acc[k] = 0.f;
  if (c0 < n) {
    for (int r = r_begin + warp; r < r_end; r += 8) {
      const uint4 raw = *reinterpret_cast<const uint4*>(x + (size_t)r * ld + c0);
      const __nv_bfloat162* h = reinterpret_cast<const __nv_bfloat162*>(&raw);
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        const float2 f = __bfloat1622float2(h[k]);
        acc[2 * k] += f.x;
        acc[2 * k + 1] += f.y;
      }
    }
  }
#pragma unroll
  for (int k = 0; k < 8; ++k) s_part[warp][lane * 8 + k] = acc[k];
  __syncthreads();
  const int c = blockIdx.x * 256 + threadIdx.x;
  if (c < n) {
    float t = 0.f;
#pragma unroll
    for (int w = 0; w < 8; ++w) t += s_part[w][threadIdx.x];
    atomicAdd(out + c, t);
  }
}

// ---- fused optimizer tail: aux gradients + optimizer step + bf16 operand refresh in two launches ------------------
// The separate passes  aux_grads (dA = s*dW, ds = <dW, A>)  ->  AdamW/SGD over the arena  ->  prep_weights
// (W_eff = W + s*A -> bf16)  collapse into
//   phase A: every trainable tensor that is not a Linear weight.  An aux_weight chunk takes its gradient on the fly,
//            s_old * dW (never materialised), and leaves its share of <dW, A> behind; the block that completes a
//            layer's last chunk also steps that layer's cross_modal_scale (nobody reads s_old of that layer any more).
//   phase B: the Linear weights.  Steps W and writes the bf16 GEMM operand bf16(W_new + s_new * A_new) from registers.
// so CrossModalReparamLinear's mix (mome.py:58-60) and its autograd never run as elementwise kernels of their own.
// Not used with gradient clipping or the FedProx term (both need every gradient in memory before the step).
struct FChunk {
  long long off;      // float offset of the chunk in the param / grad / state arenas
  int len;            // <= kChunk
  int kind;           // 0 plain | 1 aux_weight (phase A) | 2 Linear weight (phase B) | 3 Linear weight with an aux partner
  int layer;          // kinds 1, 3: index into the aux layer table
  int pad;
  long long x0;       // kind 1: offset of the matching W element;  kinds 2, 3: bf16 element offset in the operand arena
  long long x1;       // kind 3: offset of the matching aux_weight element
};
struct FAuxLayer {
  long long w_off, a_off, s_off;
  long long numel;
  int chunk_start;    // (unused here: same table as fc_aux_grads)
};
struct OptHyper {
  int opt;            // FC_OPT_ADAMW | FC_OPT_SGD
  float lr, beta1, beta2, eps, wd, bc1, bc2_sqrt;      // AdamW
  float momentum, dampening;                            // SGD
  int nesterov, first_step;
};

__device__ __forceinline__ float opt_step1(const OptHyper& h, float p, float g, float& m, float& v) {
  if (h.opt == FC_OPT_ADAMW) {
    const float w = p * (1.0f - h.lr * h.wd);
    m = m + (g - m) * (1.0f - h.beta1);
    v = v * h.beta2 + (1.0f - h.beta2) * g * g;
    const float denom = sqrtf(v) / h.bc2_sqrt + h.eps;
    return w - (h.lr / h.bc1) * (m / denom);
  }
  if (h.wd != 0.0f) g += h.wd * p;
  if (h.momentum != 0.0f) {
    m = h.first_step ? g : h.momentum * m + (1.0f - h.dampening) * g;
    g = h.nesterov ? g + h.momentum * m : m;
  }
  return p - h.lr * g;
}

template <int PHASE>
__global__ void __launch_bounds__(256) fused_opt_kernel(float* __restrict__ p, float* __restrict__ g, float* __restrict__ m,
                                                        float* __restrict__ v, __nv_bfloat16* __restrict__ ops,
                                                        const FChunk* __restrict__ chunks, int n_chunks,
                                                        const FAuxLayer* __restrict__ layers, int* __restrict__ counters,
                                                        const OptHyper h) {
  __shared__ float s_red[8];
  __shared__ int s_last;
  const bool adam = h.opt == FC_OPT_ADAMW;
  const bool use_m = adam || h.momentum != 0.0f;
  for (int c = blockIdx.x; c < n_chunks; c += gridDim.x) {
    const FChunk ch = chunks[c];
    float s_scale = 0.f;                         // kind 1: s_old (gradient of A = s_old * dW); kind 3: s_new
    if (ch.kind == 1 || ch.kind == 3) s_scale = p[layers[ch.layer].s_off];
    float dot = 0.f;
    for (int i = threadIdx.x * 4; i < ch.len; i += blockDim.x * 4) {
      const long long o = ch.off + i;
      float4 pv = ld4(p + o), gv, mv = make_float4(0.f, 0.f, 0.f, 0.f), vv = make_float4(0.f, 0.f, 0.f, 0.f);
      if (PHASE == 0 && ch.kind == 1) {          // dA = s_old * dW, and this chunk's share of ds = <dW, A>
        const float4 gw = ld4(g + ch.x0 + i);
        dot += (gw.x * pv.x + gw.y * pv.y) + (gw.z * pv.z + gw.w * pv.w);
        gv = make_float4(s_scale * gw.x, s_scale * gw.y, s_scale * gw.z, s_scale * gw.w);
      } else {
        gv = ld4(g + o);
      }
      if (use_m && !(h.opt == FC_OPT_SGD && h.first_step)) mv = ld4(m + o);
      if (adam) vv = ld4(v + o);
      float* pp = &pv.x; float* gp = &gv.x; float* mp = &mv.x; float* vp = &vv.x;
#pragma unroll
      for (int k = 0; k < 4; ++k) pp[k] = opt_step1(h, pp[k], gp[k], mp[k], vp[k]);
      st4(p + o, pv);
      if (use_m) st4(m + o, mv);
      if (adam) st4(v + o, vv);
      if (PHASE == 1) {                          // the GEMM operand of this Linear: bf16(W_new [+ s_new * A_new])
        float4 w = pv;
        if (ch.kind == 3) {
          const float4 av = ld4(p + ch.x1 + i);
          w.x += s_scale * av.x; w.y += s_scale * av.y; w.z += s_scale * av.z; w.w += s_scale * av.w;
        }
        __nv_bfloat162 lo = __floats2bfloat162_rn(w.x, w.y), hi = __floats2bfloat162_rn(w.z, w.w);
        uint2 pk;
        pk.x = *reinterpret_cast<uint32_t*>(&lo);
        pk.y = *reinterpret_cast<uint32_t*>(&hi);
        *reinterpret_cast<uint2*>(ops + ch.x0 + i) = pk;
      }
    }
    if (PHASE == 0 && ch.kind == 1) {            // CTA-uniform
      dot = warp_sum(dot);
      if ((threadIdx.x & 31) == 0) s_red[threadIdx.x >> 5] = dot;
      __syncthreads();
      if (threadIdx.x == 0) {
        float t = 0.f;
        for (int w = 0; w < 8; ++w) t += s_red[w];
        const FAuxLayer L = layers[ch.layer];
        atomicAdd(g + L.s_off, t);
        __threadfence();
        const int n_layer_chunks = (int)((L.numel + kChunk - 1) / kChunk);
        s_last = atomicAdd(counters + ch.layer, 1) == n_layer_chunks - 1;
        if (s_last) {                            // the layer's <dW, A> is complete: step its cross_modal_scale
          __threadfence();
          counters[ch.layer] = 0;                // re-armed for the next step
          const float gs = *reinterpret_cast<volatile float*>(g + L.s_off);
          float ms = use_m ? m[L.s_off] : 0.f, vs = adam ? v[L.s_off] : 0.f;
          p[L.s_off] = opt_step1(h, p[L.s_off], gs, ms, vs);
          if (use_m) m[L.s_off] = ms;
          if (adam) v[L.s_off] = vs;
        }
      }
      __syncthreads();
    }
  }
}

int grid_for(int n_items, int device, int per_sm) {
  int g = fc_num_sms(device) * per_sm;
  return n_items < g ? n_items : g;
}
// Grid of a chunk-strided kernel = the CTAs that are actually RESIDENT (occupancy of this kernel, at most `max_per_sm`):
// with 47 registers the optimizer kernels fit 5 CTAs of 256 threads per SM, so a grid of 8 per SM ran as a full wave
// of 740 followed by a 60 %-filled one.
template <typename K>
int resident_grid(K kernel, int n_items, int device, int max_per_sm) {
  int occ = 0;
  if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kernel, 256, 0) != cudaSuccess || occ < 1) occ = 1;
  return grid_for(n_items, device, occ < max_per_sm ? occ : max_per_sm);
}

}  // namespace

extern "C" int fc_chunk_floats(void) { return kChunk; }

extern "C" int fc_adamw_step(float* params, const float* grads, float* exp_avg, float* exp_avg_sq,
                             const void* chunks, int n_chunks, float lr, float beta1, float beta2, float eps,
                             float weight_decay, int step, const float* grad_sumsq, float max_norm, int device,
                             void* stream) {
  FC_REQUIRE(step >= 1, "fc_adamw_step: step must be >= 1");
  if (n_chunks <= 0) return FC_OK;
  FcDeviceGuard guard(device);
  const float bc1 = 1.0f - powf(beta1, (float)step);
  const float bc2_sqrt = sqrtf(1.0f - powf(beta2, (float)step));
  adamw_kernel<<<resident_grid(adamw_kernel, n_chunks, device, 8), 256, 0, reinterpret_cast<cudaStream_t>(stream)>>>(
      params, grads, exp_avg, exp_avg_sq, reinterpret_cast<const Chunk*>(chunks), n_chunks, lr, beta1, beta2, eps,
      weight_decay, bc1, bc2_sqrt, max_norm > 0.f ? grad_sumsq : nullptr, max_norm);
  FC_LAUNCH_CHECK();
  return FC_OK;
}

extern "C" int fc_sgd_step(float* params, const float* grads, float* momentum_buf, const void* chunks,
                           int n_chunks, float lr, float momentum, float dampening, float weight_decay, int nesterov,
                           int first_step, const float* grad_sumsq, float max_norm, int device, void* stream) {
  FC_REQUIRE(momentum == 0.0f || momentum_buf != nullptr, "fc_sgd_step: momentum needs a buffer");
  if (n_chunks <= 0) return FC_OK;
  FcDeviceGuard guard(device);
  sgd_kernel<<<grid_for(n_chunks, device, 8), 256, 0, reinterpret_cast<cudaStream_t>(stream)>>>(
      params, grads, momentum_buf, reinterpret_cast<const Chunk*>(chunks), n_chunks, lr, momentum, dampening,
      weight_decay, nesterov, first_step, max_norm > 0.f ? grad_sumsq : nullptr, max_norm);
  FC_LAUNCH_CHECK();
  return FC_OK;
}

extern "C" int fc_sumsq(const float* a, const float* b, const void* chunks, int n_chunks, float* out,
                        int single_output, int device, void* stream) {
  if (n_chunks <= 0) return FC_OK;
  FcDeviceGuard guard(device);
  sumsq_kernel<<<grid_for(n_chunks, device, 8), 256, 0, reinterpret_cast<cudaStream_t>(stream)>>>(
      a, b, reinterpret_cast<const Chunk*>(chunks), n_chunks, out, single_output);
  FC_LAUNCH_CHECK();
  return FC_OK;
}

extern "C" int fc_prox_grad(float* grads, const float* params, const float* global_params, const void* chunks,
                            int n_chunks, const float* seg_sumsq, int n_segments, float mu, float* loss_out,
                            float* loss_weighted_out, float weight, int device, void* stream) {
  if (n_chunks <= 0) return FC_OK;
  FcDeviceGuard guard(device);
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  prox_grad_kernel<<<grid_for(n_chunks, device, 8), 256, 0, st>>>(grads, params, global_params,
                                                                 reinterpret_cast<const Chunk*>(chunks), n_chunks,
                                                                 seg_sumsq, mu);
  FC_LAUNCH_CHECK();
  if (loss_out != nullptr) {
    prox_loss_kernel<<<1, 256, 0, st>>>(seg_sumsq, n_segments, mu, loss_out, loss_weighted_out, weight);
    FC_LAUNCH_CHECK();
  }
  return FC_OK;
}

extern "C" int fc_prep_weights(const float* params, void* operands_bf16, const void* layers, int n_layers,
                               int n_tiles, int device, void* stream) {
  if (n_layers <= 0 || n_tiles <= 0) return FC_OK;
  FcDeviceGuard guard(device);
  prep_weights_kernel<<<grid_for(n_tiles, device, 16), 256, 0, reinterpret_cast<cudaStream_t>(stream)>>>(
      params, reinterpret_cast<__nv_bfloat16*>(operands_bf16), reinterpret_cast<const PrepLayer*>(layers), n_layers,
      n_tiles);
  FC_LAUNCH_CHECK();
  return FC_OK;
}

extern "C" int fc_aux_grads(const float* params, float* grads, const void* layers, int n_layers, int n_chunks,
                            int aux_trained, int device, void* stream) {
  if (n_layers <= 0 || n_chunks <= 0) return FC_OK;
  FcDeviceGuard guard(device);
  aux_grads_kernel<<<grid_for(n_chunks, device, 8), 256, 0, reinterpret_cast<cudaStream_t>(stream)>>>(
      params, grads, reinterpret_cast<const AuxLayer*>(layers), n_layers, n_chunks, aux_trained);
  FC_LAUNCH_CHECK();
  return FC_OK;
}

extern "C" int fc_colsum_bf16(const void* x, long long ld, int rows, int n, float* out, int device, void* stream) {
  FC_REQUIRE(n % 8 == 0 && ld % 8 == 0, "fc_colsum_bf16: n and ld must be multiples of 8");
  if (rows <= 0 || n <= 0) return FC_OK;
  FcDeviceGuard guard(device);
  const int gx = (n + 255) / 256;
  int gy = (fc_num_sms(device) * 4 + gx - 1) / gx;
  const int max_gy = (rows + 63) / 64;
  if (gy > max_gy) gy = max_gy;
  if (gy < 1) gy = 1;
  const int rpb = (rows + gy - 1) / gy;
  colsum_bf16_kernel<<<dim3(gx, gy), 256, 0, reinterpret_cast<cudaStream_t>(stream)>>>(
      reinterpret_cast<const __nv_bfloat16*>(x), ld, rows, n, rpb, out);
  FC_LAUNCH_CHECK();
  return FC_OK;
}

// chunks_a / chunks_b: device arrays of {long long off; int len, kind, layer, pad; long long x0, x1;} (see FChunk);
// aux_layers: as for fc_aux_grads; counters: n_aux_layers zeroed ints (left zero again on return).
extern "C" int fc_opt_fused(float* params, float* grads, float* state0, float* state1, void* operands_bf16,
                            const void* chunks_a, int n_chunks_a, const void* chunks_b, int n_chunks_b,
                            const void* aux_layers, int* counters, int optimizer, float lr, float beta1, float beta2,
                            float eps, float weight_decay, float momentum, float dampening, int nesterov, int step,
                            int device, void* stream) {
  FC_REQUIRE(optimizer == FC_OPT_ADAMW || optimizer == FC_OPT_SGD, "fc_opt_fused: optimizer %d", optimizer);
  FC_REQUIRE(step >= 1, "fc_opt_fused: step must be >= 1");
  FC_REQUIRE(optimizer != FC_OPT_ADAMW || (state0 && state1), "fc_opt_fused: AdamW needs both moment arenas");
  FC_REQUIRE(optimizer != FC_OPT_SGD || momentum == 0.0f || state0, "fc_opt_fused: momentum needs a buffer");
  FcDeviceGuard guard(device);
  OptHyper h;
  h.opt = optimizer; h.lr = lr; h.beta1 = beta1; h.beta2 = beta2; h.eps = eps; h.wd = weight_decay;
  h.bc1 = 1.0f - powf(beta1, (float)step);
  h.bc2_sqrt = sqrtf(1.0f - powf(beta2, (float)step));
  h.momentum = momentum; h.dampening = dampening; h.nesterov = nesterov; h.first_step = step == 1;
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  if (n_chunks_a > 0) {
    fused_opt_kernel<0><<<resident_grid(fused_opt_kernel<0>, n_chunks_a, device, 8), 256, 0, st>>>(
        params, grads, state0, state1, reinterpret_cast<__nv_bfloat16*>(operands_bf16),
        reinterpret_cast<const FChunk*>(chunks_a), n_chunks_a, reinterpret_cast<const FAuxLayer*>(aux_layers), counters, h);
    FC_LAUNCH_CHECK();
  }
  if (n_chunks_b > 0) {
    fused_opt_kernel<1><<<resident_grid(fused_opt_kernel<1>, n_chunks_b, device, 8), 256, 0, st>>>(
        params, grads, state0, state1, reinterpret_cast<__nv_bfloat16*>(operands_bf16),
        reinterpret_cast<const FChunk*>(chunks_b), n_chunks_b, reinterpret_cast<const FAuxLayer*>(aux_layers), counters, h);
    FC_LAUNCH_CHECK();
  }
  return FC_OK;
}
