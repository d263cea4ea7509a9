// LayerNorm forward / backward over the fp32 residual stream (one warp per token row, 128-bit loads).
//
// Replaces nn.LayerNorm in Block.norm1/norm2 (eps 1e-5, /root/reference/src/models/mome.py:203,215,226-227),
// the final self.norm (eps 1e-6, :751-752) and their autograd.  The forward writes the bf16 operand the
// following tcgen05 GEMM consumes; the backward adds into the running residual gradient and also emits the
// DropPath-scaled bf16 copy that the next backward GEMM consumes, so no separate cast/scale kernel runs.
#include "common.cuh"
#include "../../include/fedcola_b200.h"

namespace {

constexpr int kMaxVec = 8;   // float4 per lane -> d <= 1024 (kernels are instantiated for NV = 1,2,3,4,6,8)

__device__ __forceinline__ uint32_t pack2(float a, float b) {
  __nv_bfloat162 v = __floats2bfloat162_rn(a, b);
  return *reinterpret_cast<uint32_t*>(&v);
}

// ---- forward -----------------------------------------------------------------------------------
template <int NV>
__global__ void __launch_bounds__(256) ln_fwd_kernel(const float* __restrict__ x, long long x_row_stride,
                                                     const float* __restrict__ gamma, const float* __restrict__ beta,
                                                     float eps, __nv_bfloat16* __restrict__ y_bf16,
                                                     float* __restrict__ y_f32, float* __restrict__ mean_out,
                                                     float* __restrict__ rstd_out, int rows, int d) {
  const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
  const int nwarps = (gridDim.x * blockDim.x) >> 5;
  const int nvec = d >> 2;
  for (int row = warp; row < rows; row += nwarps) {
    const float4* xr = reinterpret_cast<const float4*>(x + (size_t)row * x_row_stride);
    float4 v[NV];
    float s = 0.f;
#pragma unroll
    for (int i = 0; i < NV; ++i) {
      const int c = lane + i * 32;
      if (c < nvec) {
        v[i] = xr[c];
        s += (v[i].x + v[i].y) + (v[i].z + v[i].w);
      }
    }
    const float mean = warp_sum(s) / d;
    float q = 0.f;
#pragma unroll
    for (int i = 0; i < NV; ++i) {
      const int c = lane + i * 32;
      if (c < nvec) {
        const float a = v[i].x - mean, b = v[i].y - mean, e = v[i].z - mean, f = v[i].w - mean;
        q += (a * a + b * b) + (e * e + f * f);
      }
    }
    const float rstd = rsqrtf(warp_sum(q) / d + eps);
    if (lane == 0) {
      if (mean_out) mean_out[row] = mean;
      if (rstd_out) rstd_out[row] = rstd;
    }
#pragma unroll
    for (int i = 0; i < NV; ++i) {
      const int c = lane + i * 32;
      if (c < nvec) {
        const float4 g = __ldg(reinterpret_cast<const float4*>(gamma) + c);
        const float4 b = __ldg(reinterpret_cast<const float4*>(beta) + c);
        float4 o;
        o.x = (v[i].x - mean) * rstd * g.x + b.x;
        o.y = (v[i].y - mean) * rstd * g.y + b.y;
        o.z = (v[i].z - mean) * rstd * g.z + b.z;
        o.w = (v[i].w - mean) * rstd * g.w + b.w;
        if (y_bf16) reinterpret_cast<uint2*>(y_bf16 + (size_t)row * d)[c] = make_uint2(pack2(o.x, o.y), pack2(o.z, o.w));
        if (y_f32) reinterpret_cast<float4*>(y_f32 + (size_t)row * d)[c] = o;
      }
    }
  }
}

// ---- backward ----------------------------------------------------------------------------------
// dx_row = rstd * (g - mean(g) - xhat * mean(g*xhat)),  g = dy*gamma,  xhat = (x-mean)*rstd
// DY_BF16: dy is bf16 (output of a backward GEMM) else fp32.
template <int NV, bool DY_BF16>
__global__ void __launch_bounds__(256) ln_bwd_kernel(const void* __restrict__ dy_, long long dy_row_stride,
                                                     const float* __restrict__ x, long long x_row_stride,
                                                     const float* __restrict__ mean, const float* __restrict__ rstd,
                                                     const float* __restrict__ gamma, float* __restrict__ dx,
                                                     long long dx_row_stride, int accumulate,
                                                     __nv_bfloat16* __restrict__ dxs, long long dxs_row_stride,
                                                     const float* __restrict__ row_scale,
                                                     int rows_per_group, float* __restrict__ dgamma,
                                                     float* __restrict__ dbeta, float* __restrict__ dxs_colsum,
                                                     int rows, int d) {
  extern __shared__ float s_part[];   // [3][warps_per_block][d]
  const int wib = threadIdx.x >> 5, lane = threadIdx.x & 31, wpb = blockDim.x >> 5;
  const int warp = blockIdx.x * wpb + wib, nwarps = gridDim.x * wpb;
  const int nvec = d >> 2;
  float4 dg[NV], db[NV], dc[NV];
#pragma unroll
  for (int i = 0; i < NV; ++i) dg[i] = db[i] = dc[i] = make_float4(0.f, 0.f, 0.f, 0.f);

  for (int row = warp; row < rows; row += nwarps) {
    const float m = mean[row], r = rstd[row];
    const float4* xr = reinterpret_cast<const float4*>(x + (size_t)row * x_row_stride);
    float4 g[NV], xh[NV], prev[NV];
    float s1 = 0.f, s2 = 0.f;
    // every global load of the row (dy, x and the running dx) is issued before the two warp reductions: one
    // memory round trip per row instead of two
#pragma unroll
    for (int i = 0; i < NV; ++i) {
      const int c = lane + i * 32;
      prev[i] = make_float4(0.f, 0.f, 0.f, 0.f);
      if (c < nvec && accumulate) prev[i] = reinterpret_cast<const float4*>(dx + (size_t)row * dx_row_stride)[c];
    }
#pragma unroll
    for (int i = 0; i < NV; ++i) {
      const int c = lane + i * 32;
      if (c < nvec) {
        float4 dyv;
        if (DY_BF16) {
          const uint2 raw = reinterpret_cast<const uint2*>(reinterpret_cast<const __nv_bfloat16*>(dy_) +
                                                           (size_t)row * dy_row_stride)[c];
          const float2 lo = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&raw.x));
          const float2 hi = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&raw.y));
          dyv = make_float4(lo.x, lo.y, hi.x, hi.y);
        } else {
          dyv = reinterpret_cast<const float4*>(reinterpret_cast<const float*>(dy_) + (size_t)row * dy_row_stride)[c];
        }
        const float4 xv = xr[c];
        const float4 gm = __ldg(reinterpret_cast<const float4*>(gamma) + c);
        xh[i] = make_float4((xv.x - m) * r, (xv.y - m) * r, (xv.z - m) * r, (xv.w - m) * r);
        g[i] = make_float4(dyv.x * gm.x, dyv.y * gm.y, dyv.z * gm.z, dyv.w * gm.w);
        s1 += (g[i].x + g[i].y) + (g[i].z + g[i].w);
        s2 += (g[i].x * xh[i].x + g[i].y * xh[i].y) + (g[i].z * xh[i].z + g[i].w * xh[i].w);
        dg[i].x += dyv.x * xh[i].x; dg[i].y += dyv.y * xh[i].y; dg[i].z += dyv.z * xh[i].z; dg[i].w += dyv.w * xh[i].w;
        db[i].x += dyv.x; db[i].y += dyv.y; db[i].z += dyv.z; db[i].w += dyv.w;
      }
    }
    const float m1 = warp_sum(s1) / d, m2 = warp_sum(s2) / d;
    const float sc = row_scale ? __ldg(row_scale + row / rows_per_group) : 1.0f;
#pragma unroll
    for (int i = 0; i < NV; ++i) {
      const int c = lane + i * 32;
      if (c < nvec) {
        float4 o;
        o.x = r * (g[i].x - m1 - xh[i].x * m2);
        o.y = r * (g[i].y - m1 - xh[i].y * m2);
        o.z = r * (g[i].z - m1 - xh[i].z * m2);
        o.w = r * (g[i].w - m1 - xh[i].w * m2);
        float4* dxr = reinterpret_cast<float4*>(dx + (size_t)row * dx_row_stride) + c;
        o.x += prev[i].x; o.y += prev[i].y; o.z += prev[i].z; o.w += prev[i].w;
        *dxr = o;
        if (dxs) {
          const uint32_t w0 = pack2(o.x * sc, o.y * sc), w1 = pack2(o.z * sc, o.w * sc);
          reinterpret_cast<uint2*>(dxs + (size_t)row * dxs_row_stride)[c] = make_uint2(w0, w1);
          if (dxs_colsum) {     // bias gradient of the Linear that consumes dxs: sum the bf16 values it will read
            const float2 a = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&w0));
            const float2 b = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&w1));
            dc[i].x += a.x; dc[i].y += a.y; dc[i].z += b.x; dc[i].w += b.y;
          }
        }
      }
    }
  }
  if (dgamma == nullptr && dxs_colsum == nullptr) return;
  // block-level reduction of the per-warp column partials, then one atomicAdd per column per block
  float* pg = s_part;
  float* pb = s_part + wpb * d;
  float* pc = s_part + 2 * wpb * d;
#pragma unroll
  for (int i = 0; i < NV; ++i) {
    const int c = lane + i * 32;
    if (c < nvec) {
      reinterpret_cast<float4*>(pg + wib * d)[c] = dg[i];
      reinterpret_cast<float4*>(pb + wib * d)[c] = db[i];
      reinterpret_cast<float4*>(pc + wib * d)[c] = dc[i];
    }
  }
  __syncthreads();
  for (int c = threadIdx.x; c < d; c += blockDim.x) {
    float a = 0.f, b = 0.f, e = 0.f;
    for (int w = 0; w < wpb; ++w) {
      a += pg[w * d + c];
      b += pb[w * d + c];
      e += pc[w * d + c];
    }
    if (dgamma != nullptr) {
      atomicAdd(dgamma + c, a);
      atomicAdd(dbeta + c, b);
    }
    if (dxs_colsum != nullptr) atomicAdd(dxs_colsum + c, e);
  }
}

}  // namespace

extern "C" int fc_layernorm_fwd(const float* x, long long x_row_stride, const float* gamma, const float* beta,
                                float eps, void* y_bf16, float* y_f32, float* mean, float* rstd, int rows, int d,
                                int device, void* stream) {
  FC_REQUIRE(rows >= 0 && d > 0 && d % 4 == 0 && d <= kMaxVec * 128, "fc_layernorm_fwd: d=%d unsupported", d);
  FC_REQUIRE(x_row_stride % 4 == 0, "fc_layernorm_fwd: row stride must be a multiple of 4");
  if (rows == 0) return FC_OK;
  FcDeviceGuard guard(device);
  const int wpb = 8;
  int grid = (rows + wpb - 1) / wpb;
  const int cap = fc_num_sms(device) * 8;
  if (grid > cap) grid = cap;
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  auto yb = reinterpret_cast<__nv_bfloat16*>(y_bf16);
#define FC_LN_FWD(NV) ln_fwd_kernel<NV><<<grid, wpb * 32, 0, st>>>(x, x_row_stride, gamma, beta, eps, yb, y_f32, mean, rstd, rows, d)
  const int nv = (d + 127) / 128;
  if (nv <= 1) FC_LN_FWD(1); else if (nv == 2) FC_LN_FWD(2); else if (nv == 3) FC_LN_FWD(3);
  else if (nv == 4) FC_LN_FWD(4); else if (nv <= 6) FC_LN_FWD(6); else FC_LN_FWD(8);
#undef FC_LN_FWD
  FC_LAUNCH_CHECK();
  return FC_OK;
}

extern "C" int fc_layernorm_bwd(const void* dy, int dy_is_bf16, long long dy_row_stride, const float* x,
                                long long x_row_stride, const float* mean, const float* rstd, const float* gamma,
                                float* dx, long long dx_row_stride, int accumulate, void* dxs_bf16,
                                long long dxs_row_stride, const float* row_scale, int rows_per_group, float* dgamma,
                                float* dbeta, float* dxs_colsum, int rows, int d, int device, void* stream) {
  FC_REQUIRE(rows >= 0 && d > 0 && d % 4 == 0 && d <= kMaxVec * 128, "fc_layernorm_bwd: d=%d unsupported", d);
  FC_REQUIRE((dgamma == nullptr) == (dbeta == nullptr), "fc_layernorm_bwd: dgamma/dbeta must both be given");
  FC_REQUIRE(row_scale == nullptr || rows_per_group > 0, "fc_layernorm_bwd: rows_per_group");
  FC_REQUIRE(dxs_colsum == nullptr || dxs_bf16 != nullptr, "fc_layernorm_bwd: dxs_colsum needs dxs");
  if (rows == 0) return FC_OK;
  FcDeviceGuard guard(device);
  const int wpb = 8;
  int grid = (rows + wpb - 1) / wpb;
  // 2 resident CTAs/SM (111 registers); measured best of {1,2,3,4,8} per SM — also halves the per-CTA column atomics
  const int cap = fc_num_sms(device) * 2;
  if (grid > cap) grid = cap;
  const size_t smem = (dgamma || dxs_colsum) ? sizeof(float) * 3 * wpb * d : 0;
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  auto dxs = reinterpret_cast<__nv_bfloat16*>(dxs_bf16);
  const int rpg = rows_per_group > 0 ? rows_per_group : 1;
#define FC_LN_BWD(NV)                                                                                              \
  do {                                                                                                             \
    if (dy_is_bf16) {                                                                                              \
      if (smem > 48 * 1024) FC_SMEM_OPT_IN((ln_bwd_kernel<NV, true>), smem);                                       \
      ln_bwd_kernel<NV, true><<<grid, wpb * 32, smem, st>>>(dy, dy_row_stride, x, x_row_stride, mean, rstd, gamma, dx, \
                                                           dx_row_stride, accumulate, dxs, dxs_row_stride, row_scale, \
                                                           rpg, dgamma, dbeta, dxs_colsum, rows, d);               \
    } else {                                                                                                       \
      if (smem > 48 * 1024) FC_SMEM_OPT_IN((ln_bwd_kernel<NV, false>), smem);                                      \
      ln_bwd_kernel<NV, false><<<grid, wpb * 32, smem, st>>>(dy, dy_row_stride, x, x_row_stride, mean, rstd, gamma, dx, \
                                                            dx_row_stride, accumulate, dxs, dxs_row_stride, row_scale, \
                                                            rpg, dgamma, dbeta, dxs_colsum, rows, d);              \
    }                                                                                                              \
  } while (0)
  const int nv = (d + 127) / 128;
  if (nv <= 1) FC_LN_BWD(1); else if (nv == 2) FC_LN_BWD(2); else if (nv == 3) FC_LN_BWD(3);
  else if (nv == 4) FC_LN_BWD(4); else if (nv <= 6) FC_LN_BWD(6); else FC_LN_BWD(8);
#undef FC_LN_BWD
  FC_LAUNCH_CHECK();
  return FC_OK;
}
