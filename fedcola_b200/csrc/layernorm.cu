// LayerNorm forward / backward over the fp32 residual stream (one warp per token row, 128-bit loads).
//
// Replaces nn.LayerNorm in Block.norm1/norm2 (eps 1e-5, /root/reference/src/models/mome.py:203,215,226-227),
// the final self.norm (eps 1e-6, :751-752) and their autograd.  The forward writes the bf16 operand the
// following tcgen05 GEMM consumes; the backward adds into the running residual gradient and also emits the
// DropPath-scaled bf16 copy that the next backward GEMM consumes, so no separate cast/scale kernel runs.
#include "common.cuh"
#include "sm100.cuh"
#include "../../include/fedcola_b200.h"

#include <cstdlib>
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

// ---- backward, streamed through shared memory ------------------------------------------------------------
// The register version above keeps one row per warp in flight (111 registers -> 16 warps per SM): ~35 KB of loads
// outstanding per SM, 3.7 TB/s on the 66 k-row launches of a client group.  Here a producer warp streams the rows
// — dy, x and the running dx of 8 rows per stage — into a 5-stage shared-memory ring with 1-D bulk copies
// (cp.async.bulk + mbarrier transaction counts), so ~150 KB are in flight per SM whatever the consumers' register
// budget is; the 8 consumer warps (one row each per stage) read the row from shared memory, reduce, and store dx /
// dxs straight to global memory.  Same arithmetic, same order of operations per row as the register version.
constexpr int kRingRows = 8;           // rows per stage == consumer warps
constexpr int kRingThreads = (kRingRows + 1) * 32;

__device__ __forceinline__ void bulk_g2s(uint32_t dst, const void* src, uint32_t bytes, uint64_t* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
               ::"r"(dst), "l"(src), "r"(bytes), "r"(sm100::smem_u32(bar)) : "memory");
}

template <int NV, bool DY_BF16>
__global__ void __launch_bounds__(kRingThreads, (NV <= 3 ? 2 : 1))
ln_bwd_ring_kernel(const __grid_constant__ BwdSets S, long long dy_row_stride, long long x_row_stride,
                   long long dx_row_stride, int accumulate, long long dxs_row_stride, int rows_per_group, int rows,
                   int d, int stages) {
  using namespace sm100;
  extern __shared__ __align__(128) uint8_t ring[];
  const BwdSet& A = S.g[blockIdx.y];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int nvec = d >> 2;
  const int dy_bytes = d * (DY_BF16 ? 2 : 4), f_bytes = d * 4;
  const int row_bytes = dy_bytes + f_bytes + (accumulate ? f_bytes : 0);      // dy | x | dx (16-byte multiples: d % 8 == 0)
  const int stage_bytes = kRingRows * row_bytes;
  uint64_t* full = reinterpret_cast<uint64_t*>(ring + stages * stage_bytes);
  uint64_t* empty = full + stages;
  float* s_part = reinterpret_cast<float*>(ring);                              // reused after the ring has drained
  if (threadIdx.x == 0) {
    for (int s = 0; s < stages; ++s) {
      mbar_init(&full[s], 1);
      mbar_init(&empty[s], kRingRows);
    }
    fence_barrier_init();
  }
  __syncthreads();
  const int n_chunks = (rows + kRingRows - 1) / kRingRows;

  if (warp == kRingRows) {
    // ================= producer warp: lane l copies the three pieces of row l % 8 =================
    int s = 0;
    uint32_t ph = 0;
    for (int chunk = blockIdx.x; chunk < n_chunks; chunk += gridDim.x) {
      const int row0 = chunk * kRingRows;
      const int nrows = min(kRingRows, rows - row0);
      if (lane == 0) {
        mbar_wait(&empty[s], ph ^ 1);
        mbar_arrive_expect_tx(&full[s], nrows * row_bytes);
      }
      __syncwarp();
      const int r = lane & 7, piece = lane >> 3;          // piece 0: dy, 1: x, 2: running dx
      if (r < nrows && piece < (accumulate ? 3 : 2)) {
        const size_t row = row0 + r;
        const uint32_t dst = smem_u32(ring + s * stage_bytes + r * row_bytes);
        if (piece == 0) {
          const uint8_t* src = reinterpret_cast<const uint8_t*>(A.dy) + row * dy_row_stride * (DY_BF16 ? 2 : 4);
          bulk_g2s(dst, src, dy_bytes, &full[s]);
        } else if (piece == 1) {
          bulk_g2s(dst + dy_bytes, A.x + row * x_row_stride, f_bytes, &full[s]);
        } else {
          bulk_g2s(dst + dy_bytes + f_bytes, A.dx + row * dx_row_stride, f_bytes, &full[s]);
        }
      }
      if (++s == stages) { s = 0; ph ^= 1; }
    }
  } else {
    // ================= consumer warps: warp w owns row w of every stage =================
    float4 dg[NV], db[NV], dc[NV], gm[NV];
#pragma unroll
    for (int i = 0; i < NV; ++i) {
      dg[i] = db[i] = dc[i] = make_float4(0.f, 0.f, 0.f, 0.f);
      const int c = lane + i * 32;
      gm[i] = c < nvec ? __ldg(reinterpret_cast<const float4*>(A.gamma) + c) : make_float4(0.f, 0.f, 0.f, 0.f);
    }
    int s = 0;
    uint32_t ph = 0;
    for (int chunk = blockIdx.x; chunk < n_chunks; chunk += gridDim.x) {
      const int row = chunk * kRingRows + warp;
      const bool live = row < rows;                       // warp-uniform
      float m = 0.f, r = 0.f, sc = 1.0f;
      if (live) {                                         // issued before the wait: latency hides behind it
        m = __ldg(A.mean + row);
        r = __ldg(A.rstd + row);
        if (A.row_scale) sc = __ldg(A.row_scale + row / rows_per_group);
      }
      mbar_wait(&full[s], ph);
      if (live) {
        const uint8_t* base = ring + s * stage_bytes + warp * row_bytes;
        const float4* xs = reinterpret_cast<const float4*>(base + dy_bytes);
        const float4* ps = reinterpret_cast<const float4*>(base + dy_bytes + f_bytes);
        float4 g[NV], xh[NV];
        float s1 = 0.f, s2 = 0.f;
#pragma unroll
        for (int i = 0; i < NV; ++i) {
          const int c = lane + i * 32;
          if (c < nvec) {
            float4 dyv;
            if (DY_BF16) {
              const uint2 raw = reinterpret_cast<const uint2*>(base)[c];
              const float2 lo = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&raw.x));
              const float2 hi = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&raw.y));
              dyv = make_float4(lo.x, lo.y, hi.x, hi.y);
            } else {
              dyv = reinterpret_cast<const float4*>(base)[c];
            }
            const float4 xv = xs[c];
            xh[i] = make_float4((xv.x - m) * r, (xv.y - m) * r, (xv.z - m) * r, (xv.w - m) * r);
            g[i] = make_float4(dyv.x * gm[i].x, dyv.y * gm[i].y, dyv.z * gm[i].z, dyv.w * gm[i].w);
            s1 += (g[i].x + g[i].y) + (g[i].z + g[i].w);
            s2 += (g[i].x * xh[i].x + g[i].y * xh[i].y) + (g[i].z * xh[i].z + g[i].w * xh[i].w);
            dg[i].x += dyv.x * xh[i].x; dg[i].y += dyv.y * xh[i].y; dg[i].z += dyv.z * xh[i].z; dg[i].w += dyv.w * xh[i].w;
            db[i].x += dyv.x; db[i].y += dyv.y; db[i].z += dyv.z; db[i].w += dyv.w;
          }
        }
        // the two row reductions share their shuffle rounds
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
          s1 += __shfl_xor_sync(0xffffffffu, s1, o);
          s2 += __shfl_xor_sync(0xffffffffu, s2, o);
        }
        const float m1 = s1 / d, m2 = s2 / d;
#pragma unroll
        for (int i = 0; i < NV; ++i) {
          const int c = lane + i * 32;
          if (c < nvec) {
            float4 o;
            o.x = r * (g[i].x - m1 - xh[i].x * m2);
            o.y = r * (g[i].y - m1 - xh[i].y * m2);
            o.z = r * (g[i].z - m1 - xh[i].z * m2);
            o.w = r * (g[i].w - m1 - xh[i].w * m2);
            if (accumulate) {
              const float4 pv = ps[c];
              o.x += pv.x; o.y += pv.y; o.z += pv.z; o.w += pv.w;
            }
            reinterpret_cast<float4*>(A.dx + (size_t)row * dx_row_stride)[c] = o;
            if (A.dxs) {
              const uint32_t w0 = pack2(o.x * sc, o.y * sc), w1 = pack2(o.z * sc, o.w * sc);
              reinterpret_cast<uint2*>(A.dxs + (size_t)row * dxs_row_stride)[c] = make_uint2(w0, w1);
              if (A.dxs_colsum) {
                const float2 a = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&w0));
                const float2 b = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&w1));
                dc[i].x += a.x; dc[i].y += a.y; dc[i].z += b.x; dc[i].w += b.y;
              }
            }
          }
        }
      }
      __syncwarp();
      if (lane == 0) mbar_arrive(&empty[s]);              // this warp has finished reading its row of the stage
      if (++s == stages) { s = 0; ph ^= 1; }
    }
    // per-warp column partials -> shared memory (the ring is free once every consumer warp is past its last stage)
    asm volatile("bar.sync 1, %0;" ::"n"(kRingRows * 32) : "memory");
    if (A.dgamma != nullptr || A.dxs_colsum != nullptr) {
      float* pg = s_part;
      float* pb = s_part + kRingRows * d;
      float* pc = s_part + 2 * kRingRows * d;
#pragma unroll
      for (int i = 0; i < NV; ++i) {
        const int c = lane + i * 32;
        if (c < nvec) {
          reinterpret_cast<float4*>(pg + warp * d)[c] = dg[i];
          reinterpret_cast<float4*>(pb + warp * d)[c] = db[i];
          reinterpret_cast<float4*>(pc + warp * d)[c] = dc[i];
        }
      }
      asm volatile("bar.sync 1, %0;" ::"n"(kRingRows * 32) : "memory");
      for (int c = threadIdx.x; c < d; c += kRingRows * 32) {
        float a = 0.f, b = 0.f, e = 0.f;
        for (int w = 0; w < kRingRows; ++w) {
          a += pg[w * d + c];
          b += pb[w * d + c];
          e += pc[w * d + c];
        }
        if (A.dgamma != nullptr) {
          atomicAdd(A.dgamma + c, a);
          atomicAdd(A.dbeta + c, b);
        }
        if (A.dxs_colsum != nullptr) atomicAdd(A.dxs_colsum + c, e);
      }
    }
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
  const int want = (rows + wpb - 1) / wpb;
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  // grid = the CTAs that are resident for this instantiation (48 registers at d = 384: 5 CTAs of 256 threads per SM),
  // shared out over the groups and rounded down: no CTA of a second, partly filled wave
#define FC_LN_FWD(NV)                                                                                          \
  do {                                                                                                         \
    int occ = 0;                                                                                               \
    if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, ln_fwd_kernel<NV>, wpb * 32, 0) != cudaSuccess || occ < 1) occ = 1; \
    int cap = (fc_num_sms(device) * (occ < 8 ? occ : 8)) / groups;                                             \
    if (cap < 1) cap = 1;                                                                                      \
    const int grid = want < cap ? want : cap;                                                                  \
    ln_fwd_kernel<NV><<<dim3(grid, groups), wpb * 32, 0, st>>>(S, x_row_stride, eps, rows, d);                 \
  } while (0)
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
  cudaStream_t st0 = reinterpret_cast<cudaStream_t>(stream);
  {
    // streamed version: rows must be 16-byte aligned pieces (d % 8 == 0, aligned strides and bases) and the launch big
    // enough to fill the ring; FC_LN_RING=0 switches it off (measurement aid)
    static const int use_ring = getenv("FC_LN_RING") ? atoi(getenv("FC_LN_RING")) : 1;
    const int esz = dy_is_bf16 ? 2 : 4;
    bool ok = use_ring && d % 8 == 0 && (dy_row_stride * esz) % 16 == 0 && x_row_stride % 4 == 0 && dx_row_stride % 4 == 0 &&
              dxs_row_stride % 4 == 0 && (long long)rows * groups >= 4 * kRingRows * fc_num_sms(device);
    for (int g = 0; ok && g < groups; ++g)
      ok = ((reinterpret_cast<uintptr_t>(dy[g]) | reinterpret_cast<uintptr_t>(x[g]) | reinterpret_cast<uintptr_t>(dx[g]) |
             reinterpret_cast<uintptr_t>(at(dxs_bf16, g))) & 15) == 0;
    if (ok) {
      const int row_bytes = d * esz + d * 4 + (accumulate ? d * 4 : 0);
      const int stage_bytes = kRingRows * row_bytes;
      // two CTAs per SM (16 consumer warps hide the shuffle / shared-memory latencies of the per-row reductions) when
      // d <= 384 leaves room for >= 3 stages each; otherwise one CTA with the whole budget
      const int ctas_per_sm = (d <= 384 && (100 * 1024) / stage_bytes >= 3) ? 2 : 1;
      int stages = ((ctas_per_sm == 2 ? 104 : 208) * 1024) / stage_bytes;
      if (stages > 8) stages = 8;
      const int part_bytes = 3 * kRingRows * d * 4;
      if (stages >= 2 && stages * stage_bytes >= part_bytes) {
        const int smem_ring = stages * stage_bytes + 2 * stages * 8 + 64;
        const int n_chunks = (rows + kRingRows - 1) / kRingRows;
        // resident CTAs, shared out over the groups — rounded DOWN: one CTA too many would wait for a slot and run as a
        // second wave of a persistent kernel
        int gx = (fc_num_sms(device) * ctas_per_sm) / groups;
        if (gx < 1) gx = 1;
        if (gx > n_chunks) gx = n_chunks;
        const int rpg_ = rows_per_group > 0 ? rows_per_group : 1;
#define FC_LN_RING(NV)                                                                                             \
  do {                                                                                                             \
    if (dy_is_bf16) {                                                                                              \
      FC_SMEM_OPT_IN((ln_bwd_ring_kernel<NV, true>), 220 * 1024);                                                  \
      ln_bwd_ring_kernel<NV, true><<<dim3(gx, groups), kRingThreads, smem_ring, st0>>>(                            \
          S, dy_row_stride, x_row_stride, dx_row_stride, accumulate, dxs_row_stride, rpg_, rows, d, stages);       \
    } else {                                                                                                       \
      FC_SMEM_OPT_IN((ln_bwd_ring_kernel<NV, false>), 220 * 1024);                                                 \
      ln_bwd_ring_kernel<NV, false><<<dim3(gx, groups), kRingThreads, smem_ring, st0>>>(                           \
          S, dy_row_stride, x_row_stride, dx_row_stride, accumulate, dxs_row_stride, rpg_, rows, d, stages);       \
    }                                                                                                              \
  } while (0)
        const int nv_ = (d + 127) / 128;
        if (nv_ <= 1) FC_LN_RING(1); else if (nv_ == 2) FC_LN_RING(2); else if (nv_ == 3) FC_LN_RING(3);
        else if (nv_ == 4) FC_LN_RING(4); else if (nv_ <= 6) FC_LN_RING(6); else FC_LN_RING(8);
#undef FC_LN_RING
        FC_LAUNCH_CHECK();
        return FC_OK;
      }
    }
  }
  const int wpb = 8;
  int grid = (rows + wpb - 1) / wpb;
  // 2 resident CTAs/SM (111 registers); measured best of {1,2,3,4,8} per SM — also halves the per-CTA column atomics
  const int cap = (fc_num_sms(device) * 2) / groups;      // rounded down: no CTA of a second wave
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
