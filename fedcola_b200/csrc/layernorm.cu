// LayerNorm forward / backward over the fp32 residual stream (one warp per token row, 128-bit loads).
//
// Replaces nn.LayerNorm in Block.norm1/norm2 (eps 1e-5, /root/reference/src/models/mome.py:203,215,226-227),
// the final self.norm (eps 1e-6, :751-752) and their autograd.  The forward writes the bf16 operand the
// following tcgen05 GEMM consumes; the backward adds into the running residual gradient and also emits the
// DropPath-scaled bf16 copy that the next backward GEMM consumes, so no separate cast/scale kernel runs.
#include "common.cuh"
#include "../../include/fedcola_b200.h"

#include <cstring>

namespace {

constexpr int kMaxVec = 8;   // float4 per lane -> d <= 1024 (kernels are instantiated for NV = 1,2,3,4,6,8)

__device__ __forceinline__ uint32_t pack2(float a, float b) {
  __nv_bfloat162 v = __floats2bfloat162_rn(a, b);
  return *reinterpret_cast<uint32_t*>(&v);
}

// Client groups: blockIdx.y selects one of up to FC_LN_MAX_GROUPS operand sets of the same shape (the same layer of
// several clients trained in lockstep) — one launch instead of one per client.
constexpr int MAXG = FC_LN_MAX_GROUPS;
struct FwdSet {
  const float* x; const float* gamma; const float* beta;
  __nv_bfloat16* y_bf16; float* y_f32; float* mean; float* rstd;
};
struct FwdSets { FwdSet g[MAXG]; };
struct BwdSet {
  const void* dy; const float* x; const float* mean; const float* rstd; const float* gamma;
  float* dx; __nv_bfloat16* dxs; const float* row_scale; float* dgamma; float* dbeta; float* dxs_colsum;
};
struct BwdSets { BwdSet g[MAXG]; };

// ---- forward -----------------------------------------------------------------------------------
template <int NV>
__global__ void __launch_bounds__(256) ln_fwd_kernel(const __grid_constant__ FwdSets S, long long x_row_stride, float eps,
                                                     int rows, int d) {
  const FwdSet& A = S.g[blockIdx.y];
  const float* __restrict__ x = A.x;
  const float* __restrict__ gamma = A.gamma;
  const float* __restrict__ beta = A.beta;
  __nv_bfloat16* __restrict__ y_bf16 = A.y_bf16;
  float* __restrict__ y_f32 = A.y_f32;
  float* __restrict__ mean_out = A.mean;
  float* __restrict__ rstd_out = A.rstd;
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
__global__ void __launch_bounds__(256) ln_bwd_kernel(const __grid_constant__ BwdSets S, long long dy_row_stride,
                                                     long long x_row_stride, long long dx_row_stride, int accumulate,
                                                     long long dxs_row_stride, int rows_per_group, int rows, int d) {
  const BwdSet& A = S.g[blockIdx.y];
  const void* __restrict__ dy_ = A.dy;
  const float* __restrict__ x = A.x;
  const float* __restrict__ mean = A.mean;
  const float* __restrict__ rstd = A.rstd;
  const float* __restrict__ gamma = A.gamma;
  float* __restrict__ dx = A.dx;
  __nv_bfloat16* __restrict__ dxs = A.dxs;
  const float* __restrict__ row_scale = A.row_scale;
  float* __restrict__ dgamma = A.dgamma;
  float* __restrict__ dbeta = A.dbeta;
  float* __restrict__ dxs_colsum = A.dxs_colsum;
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

extern "C" int fc_layernorm_fwd_grouped(int groups, const float* const* x, long long x_row_stride,
                                        const float* const* gamma, const float* const* beta, float eps,
                                        void* const* y_bf16, float* const* y_f32, float* const* mean, float* const* rstd,
                                        int rows, int d, int device, void* stream) {
  FC_REQUIRE(groups >= 1 && groups <= MAXG, "fc_layernorm_fwd: %d groups (1..%d)", groups, MAXG);
  FC_REQUIRE(rows >= 0 && d > 0 && d % 4 == 0 && d <= kMaxVec * 128, "fc_layernorm_fwd: d=%d unsupported", d);
  FC_REQUIRE(x_row_stride % 4 == 0, "fc_layernorm_fwd: row stride must be a multiple of 4");
  if (rows == 0) return FC_OK;
  FcDeviceGuard guard(device);
  FwdSets S;
  memset(&S, 0, sizeof(S));
  for (int g = 0; g < groups; ++g) {
    FC_REQUIRE(x[g] && gamma[g] && beta[g], "fc_layernorm_fwd: null operand (group %d)", g);
    S.g[g] = FwdSet{x[g], gamma[g], beta[g], y_bf16 ? reinterpret_cast<__nv_bfloat16*>(y_bf16[g]) : nullptr,
                    y_f32 ? y_f32[g] : nullptr, mean ? mean[g] : nullptr, rstd ? rstd[g] : nullptr};
  }
  const int wpb = 8;
  int grid = (rows + wpb - 1) / wpb;
  const int cap = (fc_num_sms(device) * 8 + groups - 1) / groups;
  if (grid > cap) grid = cap;
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
#define FC_LN_FWD(NV) ln_fwd_kernel<NV><<<dim3(grid, groups), wpb * 32, 0, st>>>(S, x_row_stride, eps, rows, d)
  const int nv = (d + 127) / 128;
  if (nv <= 1) FC_LN_FWD(1); else if (nv == 2) FC_LN_FWD(2); else if (nv == 3) FC_LN_FWD(3);
  else if (nv == 4) FC_LN_FWD(4); else if (nv <= 6) FC_LN_FWD(6); else FC_LN_FWD(8);
#undef FC_LN_FWD
  FC_LAUNCH_CHECK();
  return FC_OK;
}

extern "C" int fc_layernorm_fwd(const float* x, long long x_row_stride, const float* gamma, const float* beta,
                                float eps, void* y_bf16, float* y_f32, float* mean, float* rstd, int rows, int d,
                                int device, void* stream) {
  return fc_layernorm_fwd_grouped(1, &x, x_row_stride, &gamma, &beta, eps, &y_bf16, &y_f32, &mean, &rstd, rows, d, device,
                                  stream);
}

extern "C" int fc_layernorm_bwd_grouped(int groups, const void* const* dy, int dy_is_bf16, long long dy_row_stride,
                                        const float* const* x, long long x_row_stride, const float* const* mean,
                                        const float* const* rstd, const float* const* gamma, float* const* dx,
                                        long long dx_row_stride, int accumulate, void* const* dxs_bf16,
                                        long long dxs_row_stride, const float* const* row_scale, int rows_per_group,
                                        float* const* dgamma, float* const* dbeta, float* const* dxs_colsum, int rows,
                                        int d, int device, void* stream) {
  FC_REQUIRE(groups >= 1 && groups <= MAXG, "fc_layernorm_bwd: %d groups (1..%d)", groups, MAXG);
  FC_REQUIRE(rows >= 0 && d > 0 && d % 4 == 0 && d <= kMaxVec * 128, "fc_layernorm_bwd: d=%d unsupported", d);
  auto at = [](auto tbl, int g) { return tbl ? tbl[g] : nullptr; };
  FC_REQUIRE((at(dgamma, 0) == nullptr) == (at(dbeta, 0) == nullptr), "fc_layernorm_bwd: dgamma/dbeta must both be given");
  FC_REQUIRE(at(row_scale, 0) == nullptr || rows_per_group > 0, "fc_layernorm_bwd: rows_per_group");
  FC_REQUIRE(at(dxs_colsum, 0) == nullptr || at(dxs_bf16, 0) != nullptr, "fc_layernorm_bwd: dxs_colsum needs dxs");
  if (rows == 0) return FC_OK;
  FcDeviceGuard guard(device);
  BwdSets S;
  memset(&S, 0, sizeof(S));
  for (int g = 0; g < groups; ++g) {
    FC_REQUIRE(dy[g] && x[g] && mean[g] && rstd[g] && gamma[g] && dx[g], "fc_layernorm_bwd: null operand (group %d)", g);
    FC_REQUIRE((at(dgamma, g) == nullptr) == (at(dgamma, 0) == nullptr) && (at(dxs_colsum, g) == nullptr) == (at(dxs_colsum, 0) == nullptr) &&
               (at(dxs_bf16, g) == nullptr) == (at(dxs_bf16, 0) == nullptr) && (at(row_scale, g) == nullptr) == (at(row_scale, 0) == nullptr),
               "fc_layernorm_bwd: groups disagree on optional operands");
    S.g[g] = BwdSet{dy[g], x[g], mean[g], rstd[g], gamma[g], dx[g], reinterpret_cast<__nv_bfloat16*>(at(dxs_bf16, g)),
                    at(row_scale, g), at(dgamma, g), at(dbeta, g), at(dxs_colsum, g)};
  }
  const int wpb = 8;
  int grid = (rows + wpb - 1) / wpb;
  // 2 resident CTAs/SM (111 registers); measured best of {1,2,3,4,8} per SM — also halves the per-CTA column atomics
  const int cap = (fc_num_sms(device) * 2 + groups - 1) / groups;
  if (grid > cap) grid = cap;
  const size_t smem = (at(dgamma, 0) || at(dxs_colsum, 0)) ? sizeof(float) * 3 * wpb * d : 0;
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  const int rpg = rows_per_group > 0 ? rows_per_group : 1;
#define FC_LN_BWD(NV)                                                                                              \
  do {                                                                                                             \
    if (dy_is_bf16) {                                                                                              \
      if (smem > 48 * 1024) FC_SMEM_OPT_IN((ln_bwd_kernel<NV, true>), smem);                                       \
      ln_bwd_kernel<NV, true><<<dim3(grid, groups), wpb * 32, smem, st>>>(S, dy_row_stride, x_row_stride, dx_row_stride, \
                                                                         accumulate, dxs_row_stride, rpg, rows, d); \
    } else {                                                                                                       \
      if (smem > 48 * 1024) FC_SMEM_OPT_IN((ln_bwd_kernel<NV, false>), smem);                                      \
      ln_bwd_kernel<NV, false><<<dim3(grid, groups), wpb * 32, smem, st>>>(S, dy_row_stride, x_row_stride, dx_row_stride, \
                                                                          accumulate, dxs_row_stride, rpg, rows, d); \
    }                                                                                                              \
  } while (0)
  const int nv = (d + 127) / 128;
  if (nv <= 1) FC_LN_BWD(1); else if (nv == 2) FC_LN_BWD(2); else if (nv == 3) FC_LN_BWD(3);
  else if (nv == 4) FC_LN_BWD(4); else if (nv <= 6) FC_LN_BWD(6); else FC_LN_BWD(8);
#undef FC_LN_BWD
  FC_LAUNCH_CHECK();
  return FC_OK;
}

extern "C" int fc_layernorm_bwd(const void* dy, int dy_is_bf16, long long dy_row_stride, const float* x,
                                long long x_row_stride, const float* mean, const float* rstd, const float* gamma,
                                float* dx, long long dx_row_stride, int accumulate, void* dxs_bf16,
                                long long dxs_row_stride, const float* row_scale, int rows_per_group, float* dgamma,
                                float* dbeta, float* dxs_colsum, int rows, int d, int device, void* stream) {
  return fc_layernorm_bwd_grouped(1, &dy, dy_is_bf16, dy_row_stride, &x, x_row_stride, &mean, &rstd, &gamma, &dx,
                                  dx_row_stride, accumulate, &dxs_bf16, dxs_row_stride, &row_scale, rows_per_group, &dgamma,
                                  &dbeta, &dxs_colsum, rows, d, device, stream);
}
