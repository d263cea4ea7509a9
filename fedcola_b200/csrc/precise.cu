// Kernels of the fp32-accurate validation mode (precision = 'fp32').
//
// The reference computes everything in strict fp32 (/root/reference/src/models/mome.py:150-168: fp32 QK^T and softmax;
// F.linear with TF32 off).  The production path rounds GEMM / attention operands to bf16 (tolerance 2e-2); this mode
// keeps every activation in fp32 and feeds the tensor cores SPLIT operands — X = X_hi + X_lo as two bf16 matrices,
// three tcgen05 passes per product (fc_gemm_split) — so logits, losses and gradients agree with the reference to 1e-4.
// Nothing here is tuned: it is the yardstick the fast path is held against, and it must be simple enough to trust.
//   fc_split_bf16        fp32 -> (hi, lo) bf16 pair, optionally scaled per row group (DropPath)
//   fc_prep_weights_split  W_eff = W + s*A  -> (hi, lo)                     ref: mome.py:58-60
//   fc_gelu_f32_fwd/bwd  exact-erf GELU and its derivative                  ref: mome.py:112-119 (nn.GELU())
//   fc_colsum_f32        bias gradients (column sums)
//   fc_attention_f32_fwd/bwd  softmax(q k^T / 8) v, one CTA per (sample, head), fp32 FMA     ref: mome.py:153-165
//   fc_im2col16_f32 / fc_drop_cls_rows   patch matrices in fp32             ref: mome.py:252-266
#include "common.cuh"

namespace {

__device__ __forceinline__ void split1(float v, __nv_bfloat16& hi, __nv_bfloat16& lo) {
  hi = __float2bfloat16_rn(v);
  lo = __float2bfloat16_rn(v - __bfloat162float(hi));
}

__global__ void __launch_bounds__(256) split_kernel(const float* __restrict__ x, __nv_bfloat16* __restrict__ hi,
                                                    __nv_bfloat16* __restrict__ lo, long long n, int row_len,
                                                    const float* __restrict__ row_scale, int rows_per_group) {
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    float v = x[i];
    if (row_scale != nullptr) v *= row_scale[(i / row_len) / rows_per_group];
    split1(v, hi[i], lo[i]);
  }
}

struct PrepLayer {                     // same table as fc_prep_weights (optim.cu)
  long long w_off, a_off, s_off;
  long long dst_off, dstT_off;
  int rows, cols;
  int tile_start;
};

__global__ void __launch_bounds__(256) prep_split_kernel(const float* __restrict__ params, __nv_bfloat16* __restrict__ hi,
                                                         __nv_bfloat16* __restrict__ lo, const PrepLayer* __restrict__ layers,
                                                         int n_layers) {
  for (int l = blockIdx.y; l < n_layers; l += gridDim.y) {
    const PrepLayer L = layers[l];
    const long long n = (long long)L.rows * L.cols;
    const float s = L.a_off >= 0 ? params[L.s_off] : 0.f;
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
      float w = params[L.w_off + i];
      if (L.a_off >= 0) w = w + s * params[L.a_off + i];         // weight + cross_modal_scale * aux_weight (mome.py:59)
      split1(w, hi[L.dst_off + i], lo[L.dst_off + i]);
    }
  }
}

__global__ void __launch_bounds__(256) gelu_fwd_kernel(const float* __restrict__ pre, float* __restrict__ act, long long n) {
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    const float x = pre[i];
    act[i] = 0.5f * x * (1.0f + erff(x * 0.70710678118654752440f));
  }
}
__global__ void __launch_bounds__(256) gelu_bwd_kernel(const float* __restrict__ d_act, const float* __restrict__ pre,
                                                       float* __restrict__ d_pre, long long n) {
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    const float x = pre[i];
    const float cdf = 0.5f * (1.0f + erff(x * 0.70710678118654752440f));
    const float pdf = 0.39894228040143267794f * expf(-0.5f * x * x);
    d_pre[i] = d_act[i] * (cdf + x * pdf);
  }
}

// out[c] += sum_r scale(r) * x[r, c]; one thread per column, rows strided over blockIdx.y
__global__ void __launch_bounds__(256) colsum_f32_kernel(const float* __restrict__ x, long long ld, int rows, int n,
                                                         const float* __restrict__ row_scale, int rows_per_group,
                                                         float* __restrict__ out) {
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= n) return;
  float acc = 0.f;
  for (int r = blockIdx.y; r < rows; r += gridDim.y) {
    float v = x[(size_t)r * ld + c];
    if (row_scale != nullptr) v *= row_scale[r / rows_per_group];
    acc += v;
  }
  atomicAdd(out + c, acc);
}

// ---- attention, fp32, one CTA per (sample, head); K and V of the head in shared memory -----------------------
constexpr int HD = 64;
__global__ void __launch_bounds__(128) attn_f32_fwd_kernel(const float* __restrict__ qkv, float* __restrict__ out,
                                                           float* __restrict__ lse, int N, int H) {
  extern __shared__ float sm[];
  float* sK = sm;
  float* sV = sm + (size_t)N * HD;
  const int b = blockIdx.x / H, h = blockIdx.x % H, d3 = 3 * H * HD;
  const float* base = qkv + (size_t)b * N * d3 + h * HD;
  for (int i = threadIdx.x; i < N * HD; i += blockDim.x) {
    const int t = i / HD, c = i % HD;
    sK[i] = base[(size_t)t * d3 + H * HD + c];
    sV[i] = base[(size_t)t * d3 + 2 * H * HD + c];
  }
  __syncthreads();
  for (int q = threadIdx.x; q < N; q += blockDim.x) {
    float qv[HD];
#pragma unroll
    for (int c = 0; c < HD; ++c) qv[c] = base[(size_t)q * d3 + c] * 0.125f;       // q * head_dim^-0.5 (mome.py:156)
    float mx = -INFINITY;
    for (int j = 0; j < N; ++j) {
      float s = 0.f;
#pragma unroll
      for (int c = 0; c < HD; ++c) s = fmaf(qv[c], sK[j * HD + c], s);
      mx = fmaxf(mx, s);
    }
    float o[HD];
#pragma unroll
    for (int c = 0; c < HD; ++c) o[c] = 0.f;
    float l = 0.f;
    for (int j = 0; j < N; ++j) {
      float s = 0.f;
#pragma unroll
      for (int c = 0; c < HD; ++c) s = fmaf(qv[c], sK[j * HD + c], s);
      const float p = expf(s - mx);
      l += p;
#pragma unroll
      for (int c = 0; c < HD; ++c) o[c] = fmaf(p, sV[j * HD + c], o[c]);
    }
    const float inv = 1.0f / l;
    float* dst = out + ((size_t)b * N + q) * H * HD + h * HD;
#pragma unroll
    for (int c = 0; c < HD; ++c) dst[c] = o[c] * inv;
    lse[((size_t)b * H + h) * N + q] = mx + logf(l);
  }
}

// dqkv must be zero-initialised by the caller (the K / V thirds are accumulated with atomics)
__global__ void __launch_bounds__(128) attn_f32_bwd_kernel(const float* __restrict__ qkv, const float* __restrict__ out,
                                                           const float* __restrict__ d_out, const float* __restrict__ lse,
                                                           float* __restrict__ dqkv, int N, int H) {
  extern __shared__ float sm[];
  float* sK = sm;
  float* sV = sm + (size_t)N * HD;
  float* sdK = sm + 2 * (size_t)N * HD;
  float* sdV = sm + 3 * (size_t)N * HD;
  const int b = blockIdx.x / H, h = blockIdx.x % H, d3 = 3 * H * HD, d = H * HD;
  const float* base = qkv + (size_t)b * N * d3 + h * HD;
  for (int i = threadIdx.x; i < N * HD; i += blockDim.x) {
    const int t = i / HD, c = i % HD;
    sK[i] = base[(size_t)t * d3 + d + c];
    sV[i] = base[(size_t)t * d3 + 2 * d + c];
    sdK[i] = 0.f;
    sdV[i] = 0.f;
  }
  __syncthreads();
  for (int q = threadIdx.x; q < N; q += blockDim.x) {
    float qv[HD], go[HD], dq[HD];
    const float* orow = out + ((size_t)b * N + q) * d + h * HD;
    const float* grow = d_out + ((size_t)b * N + q) * d + h * HD;
    float D = 0.f;
#pragma unroll
    for (int c = 0; c < HD; ++c) {
      qv[c] = base[(size_t)q * d3 + c] * 0.125f;
      go[c] = grow[c];
      D = fmaf(go[c], orow[c], D);
      dq[c] = 0.f;
    }
    const float L = lse[((size_t)b * H + h) * N + q];
    for (int t = 0; t < N; ++t) {
      const int j = (q + t) % N;           // staggered start: threads of a warp hit different keys' accumulators
      float s = 0.f, dp = 0.f;
#pragma unroll
      for (int c = 0; c < HD; ++c) {
        s = fmaf(qv[c], sK[j * HD + c], s);
        dp = fmaf(go[c], sV[j * HD + c], dp);
      }
      const float p = expf(s - L);
      const float ds = p * (dp - D);
#pragma unroll
      for (int c = 0; c < HD; ++c) {
        dq[c] = fmaf(ds, sK[j * HD + c], dq[c]);
        atomicAdd(&sdK[j * HD + c], ds * qv[c]);           // (qv already carries the 1/8)
        atomicAdd(&sdV[j * HD + c], p * go[c]);
      }
    }
    float* dst = dqkv + ((size_t)b * N + q) * d3 + h * HD;
#pragma unroll
    for (int c = 0; c < HD; ++c) dst[c] = dq[c] * 0.125f;
  }
  __syncthreads();
  float* dbase = dqkv + (size_t)b * N * d3 + h * HD;
  for (int i = threadIdx.x; i < N * HD; i += blockDim.x) {
    const int t = i / HD, c = i % HD;
    dbase[(size_t)t * d3 + d + c] = sdK[i];
    dbase[(size_t)t * d3 + 2 * d + c] = sdV[i];
  }
}

// img fp32 [B, C, HW, HW] -> patches fp32 [B*P, 768] (K index = c*256 + ph*16 + pw); in_chans 1 is repeated to 3
__global__ void __launch_bounds__(256) im2col16_f32_kernel(const float* __restrict__ img, float* __restrict__ patches, int B,
                                                           int Cin, int HW, int grid_p) {
  const int P = grid_p * grid_p;
  const long long total = (long long)B * P * 768;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int k = (int)(i % 768);
    const long long row = i / 768;
    const int c = k >> 8, ph = (k >> 4) & 15, pw = k & 15;
    const int b = (int)(row / P), t = (int)(row % P);
    const int py = t / grid_p, px = t % grid_p;
    const int cs = Cin == 1 ? 0 : c;
    patches[i] = img[(((size_t)b * Cin + cs) * HW + (py * 16 + ph)) * HW + px * 16 + pw];
  }
}
// dx fp32 [B, P+1, d] -> rows 1..P of every sample, fp32 [B*P, d]
__global__ void __launch_bounds__(256) drop_cls_rows_kernel(const float* __restrict__ dx, float* __restrict__ dxp, int B,
                                                            int P, int d) {
  const long long total = (long long)B * P * d;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int c = (int)(i % d);
    const long long row = i / d;
    const int b = (int)(row / P), t = (int)(row % P);
    dxp[i] = dx[((size_t)b * (P + 1) + 1 + t) * d + c];
  }
}

int grid1d(long long n, int device) {
  long long g = (n + 255) / 256;
  const long long cap = (long long)fc_num_sms(device) * 16;
  return (int)(g < cap ? (g > 0 ? g : 1) : cap);
}

}  // namespace

extern "C" int fc_split_bf16(const float* x, void* hi, void* lo, long long n, int row_len, const float* row_scale,
                             int rows_per_group, int device, void* stream) {
  if (n <= 0) return FC_OK;
  FC_REQUIRE(x && hi && lo, "fc_split_bf16: null pointer");
  FC_REQUIRE(row_scale == nullptr || (row_len > 0 && rows_per_group > 0), "fc_split_bf16: row geometry");
  FcDeviceGuard guard(device);
  split_kernel<<<grid1d(n, device), 256, 0, reinterpret_cast<cudaStream_t>(stream)>>>(
      x, reinterpret_cast<__nv_bfloat16*>(hi), reinterpret_cast<__nv_bfloat16*>(lo), n, row_len > 0 ? row_len : 1, row_scale,
      rows_per_group > 0 ? rows_per_group : 1);
  FC_LAUNCH_CHECK();
  return FC_OK;
}

extern "C" int fc_prep_weights_split(const float* params, void* hi, void* lo, const void* layers, int n_layers, int device,
                                     void* stream) {
  if (n_layers <= 0) return FC_OK;
  FcDeviceGuard guard(device);
  prep_split_kernel<<<dim3(64, n_layers < 64 ? n_layers : 64), 256, 0, reinterpret_cast<cudaStream_t>(stream)>>>(
      params, reinterpret_cast<__nv_bfloat16*>(hi), reinterpret_cast<__nv_bfloat16*>(lo),
      reinterpret_cast<const PrepLayer*>(layers), n_layers);
  FC_LAUNCH_CHECK();
  return FC_OK;
}

extern "C" int fc_gelu_f32_fwd(const float* pre, float* act, long long n, int device, void* stream) {
  if (n <= 0) return FC_OK;
  FcDeviceGuard guard(device);
  gelu_fwd_kernel<<<grid1d(n, device), 256, 0, reinterpret_cast<cudaStream_t>(stream)>>>(pre, act, n);
  FC_LAUNCH_CHECK();
  return FC_OK;
}
extern "C" int fc_gelu_f32_bwd(const float* d_act, const float* pre, float* d_pre, long long n, int device, void* stream) {
  if (n <= 0) return FC_OK;
  FcDeviceGuard guard(device);
  gelu_bwd_kernel<<<grid1d(n, device), 256, 0, reinterpret_cast<cudaStream_t>(stream)>>>(d_act, pre, d_pre, n);
  FC_LAUNCH_CHECK();
  return FC_OK;
}

extern "C" int fc_colsum_f32(const float* x, long long ld, int rows, int n, const float* row_scale, int rows_per_group,
                             float* out, int device, void* stream) {
  if (rows <= 0 || n <= 0) return FC_OK;
  FcDeviceGuard guard(device);
  int gy = rows < 64 ? rows : 64;
  colsum_f32_kernel<<<dim3((n + 255) / 256, gy), 256, 0, reinterpret_cast<cudaStream_t>(stream)>>>(
      x, ld, rows, n, row_scale, rows_per_group > 0 ? rows_per_group : 1, out);
  FC_LAUNCH_CHECK();
  return FC_OK;
}

extern "C" int fc_attention_f32_fwd(const float* qkv, float* out, float* lse, int B, int N, int H, int device, void* stream) {
  FC_REQUIRE(B > 0 && N > 0 && N <= 440 && H > 0, "fc_attention_f32_fwd: unsupported shape B=%d N=%d H=%d", B, N, H);
  FcDeviceGuard guard(device);
  const int smem = 2 * N * HD * 4;
  FC_SMEM_OPT_IN(attn_f32_fwd_kernel, 227 * 1024);
  attn_f32_fwd_kernel<<<B * H, 128, smem, reinterpret_cast<cudaStream_t>(stream)>>>(qkv, out, lse, N, H);
  FC_LAUNCH_CHECK();
  return FC_OK;
}
extern "C" int fc_attention_f32_bwd(const float* qkv, const float* out, const float* d_out, const float* lse, float* dqkv,
                                    int B, int N, int H, int device, void* stream) {
  FC_REQUIRE(B > 0 && N > 0 && N <= 220 && H > 0, "fc_attention_f32_bwd: unsupported shape B=%d N=%d H=%d (N <= 220)", B, N, H);
  FcDeviceGuard guard(device);
  const int smem = 4 * N * HD * 4;
  FC_SMEM_OPT_IN(attn_f32_bwd_kernel, 227 * 1024);
  attn_f32_bwd_kernel<<<B * H, 128, smem, reinterpret_cast<cudaStream_t>(stream)>>>(qkv, out, d_out, lse, dqkv, N, H);
  FC_LAUNCH_CHECK();
  return FC_OK;
}

extern "C" int fc_im2col16_f32(const float* img, float* patches, int B, int in_chans, int img_size, int device, void* stream) {
  FC_REQUIRE(img_size % 16 == 0 && (in_chans == 3 || in_chans == 1), "fc_im2col16_f32: img_size %% 16, in_chans in {1,3}");
  if (B <= 0) return FC_OK;
  FcDeviceGuard guard(device);
  const int gp = img_size / 16;
  im2col16_f32_kernel<<<grid1d((long long)B * gp * gp * 768, device), 256, 0, reinterpret_cast<cudaStream_t>(stream)>>>(
      img, patches, B, in_chans, img_size, gp);
  FC_LAUNCH_CHECK();
  return FC_OK;
}
extern "C" int fc_drop_cls_rows(const float* dx, float* dxp, int B, int patches, int d, int device, void* stream) {
  if (B <= 0) return FC_OK;
  FcDeviceGuard guard(device);
  drop_cls_rows_kernel<<<grid1d((long long)B * patches * d, device), 256, 0, reinterpret_cast<cudaStream_t>(stream)>>>(
      dx, dxp, B, patches, d);
  FC_LAUNCH_CHECK();
  return FC_OK;
}
