// Server aggregation: one streaming kernel over flat fp32 parameter arenas.
//
// Replaces the accumulation stage of FedavgServer._aggregate
// (/root/reference/src/server/fedavgserver.py:597,656-666) together with the aux merge that
// FedavgClient.upload performs on every call (/root/reference/src/client/fedavgclient.py:158-184).
//
// Semantics (mode FC_AGG_LERP, bit-exact with the reference on one GPU):
//   for every output segment (global model g, parameter p):  f <- old global
//     for contributing clients k in ascending id:   l = W_k (+ A_k * s_k when the client uploads an
//     aux-merged weight);  f <- f + ((l - f) * c[g,p,k])        -- three separately rounded fp32 ops
//   new global <- f
// Mode FC_AGG_WSUM evaluates the closed form  f = w_g*g + sum_k w_k*l_k  (weights computed in fp64 on
// the host) and is what the multi-GPU path uses: every rank reduces its own clients, then one NCCL
// all-reduce over the concatenated global arenas finishes the sum.
//
// Data movement: a *job* is one parameter name; it owns up to FC_AGG_MAX_OUT outputs (the globals that
// hold that name) and the list of clients that contribute to at least one of them.  Each client value is
// loaded ONCE (128-bit, L1-bypassing) and folded into all outputs, so HBM traffic equals the algorithmic
// bytes of SURVEY.md §8d:  4*[sum over contributing client tensors numel*(1+aux) + 2*numel per output].
#include "common.cuh"
#include "../../include/fedcola_b200.h"

namespace {

constexpr int kThreads = 256;
constexpr int kTileFloats = kThreads * 4;     // one float4 per thread: 1024 floats = 4 KB per source per tile
constexpr int kStage = 64;                    // source entries staged in smem per pass
constexpr int kUnroll = 4;                    // source loads in flight per thread (x 16 B)
constexpr int kRun = 8;                       // consecutive tiles a CTA takes at a time

struct AggParams {
  int mode, n_jobs, n_tiles;
  const int* job_tile_start;          // [n_jobs+1]
  const long long* job_numel;         // [n_jobs]
  const int* job_nout;                // [n_jobs]
  const unsigned long long* job_gin;  // [n_jobs*MAX_OUT]
  const unsigned long long* job_gout; // [n_jobs*MAX_OUT]
  const float* job_gscale;            // [n_jobs*MAX_OUT]   (WSUM: weight of the old global; LERP: unused)
  const int* job_src_start;           // [n_jobs+1]
  const unsigned long long* src_ptr;  // [nnz]
  const int* src_flag;                // [nnz] FC_AGG_SRC_*
  const unsigned long long* scale_ptr;// [nnz] (MERGE entries: address of the 1-element cross_modal_scale)
  const float* coef;                  // [nnz*MAX_OUT]
};

template <int MODE>
__device__ __forceinline__ float fold1(float f, float l, float c) {
  if (MODE == FC_AGG_LERP) {
    // (local - final) * c, then final += ...; three roundings, no FMA contraction (= torch CPU fp32)
    return __fadd_rn(f, __fmul_rn(__fsub_rn(l, f), c));
  } else {
    return fmaf(c, l, f);
  }
}

template <int MODE>
__global__ void __launch_bounds__(kThreads, 4) aggregate_kernel(const AggParams P) {
  __shared__ const float* s_src[kStage];
  __shared__ int s_flag[kStage];
  __shared__ float s_scale[kStage];
  __shared__ float s_coef[kStage][FC_AGG_MAX_OUT];
  __shared__ int s_job;

  // A CTA takes runs of kRun consecutive tiles: consecutive tiles almost always belong to the same job, so the
  // tile -> job search, the job's metadata and its staged source table are reused instead of being re-fetched
  // through three dependent global-memory round trips per 4 KB tile (that chain, not bandwidth, bounded rounds
  // with few clients per tensor).
  const int n_runs = (P.n_tiles + kRun - 1) / kRun;
  int job = -1, job_first = 0, job_end = 0, nout = 0, s0 = 0, s1 = 0;
  long long numel = 0;
  bool staged = false;                        // the smem table holds ALL sources of `job`
  for (int run = blockIdx.x; run < n_runs; run += gridDim.x) {
   const int tile_end = min(P.n_tiles, (run + 1) * kRun);
   for (int tile = run * kRun; tile < tile_end; ++tile) {
    if (job < 0 || tile >= job_end) {         // CTA-uniform
      __syncthreads();                        // everyone is done with s_job and the staged table
      // tile -> job: binary search over the tile prefix (n_jobs is a few hundred)
      if (threadIdx.x == 0) {
        int lo = 0, hi = P.n_jobs;            // invariant: start[lo] <= tile < start[hi]
        while (hi - lo > 1) {
          int mid = (lo + hi) >> 1;
          if (__ldg(P.job_tile_start + mid) <= tile) lo = mid; else hi = mid;
        }
        s_job = lo;
      }
      __syncthreads();
      job = s_job;
      nout = P.job_nout[job];
      numel = (P.job_numel[job] + 3) & ~3LL;  // segments are padded to 32 floats
      job_first = P.job_tile_start[job];
      job_end = P.job_tile_start[job + 1];
      s0 = P.job_src_start[job];
      s1 = P.job_src_start[job + 1];
      staged = false;
    }
    const long long idx = (long long)(tile - job_first) * kTileFloats + threadIdx.x * 4;
    const bool active = idx < numel;

    float4 f[FC_AGG_MAX_OUT];
#pragma unroll
    for (int o = 0; o < FC_AGG_MAX_OUT; ++o) {
      f[o] = make_float4(0.f, 0.f, 0.f, 0.f);
      if (active && o < nout) {
        const float* g = reinterpret_cast<const float*>(P.job_gin[job * FC_AGG_MAX_OUT + o]);
        if (MODE == FC_AGG_LERP) {
          f[o] = ld_stream_f4(g + idx);
        } else {
          const float wg = P.job_gscale[job * FC_AGG_MAX_OUT + o];
          if (wg != 0.0f) {
            const float4 x = ld_stream_f4(g + idx);
            f[o] = make_float4(wg * x.x, wg * x.y, wg * x.z, wg * x.w);
          }
        }
      }
    }
    float4 pend = make_float4(0.f, 0.f, 0.f, 0.f);    // W of an aux-merged upload, waiting for its A entry

    for (int k0 = s0; k0 < s1; k0 += kStage) {
      const int kn = min(kStage, s1 - k0);
      if (!staged) {
      __syncthreads();   // previous pass finished reading the staged table
      if (threadIdx.x < kn) {
        const int k = k0 + threadIdx.x;
        s_src[threadIdx.x] = reinterpret_cast<const float*>(P.src_ptr[k]);
        s_flag[threadIdx.x] = P.src_flag[k];
        const float* sp = reinterpret_cast<const float*>(P.scale_ptr[k]);
        s_scale[threadIdx.x] = sp ? __ldg(sp) : 0.0f;
#pragma unroll
        for (int o = 0; o < FC_AGG_MAX_OUT; ++o) s_coef[threadIdx.x][o] = P.coef[(size_t)k * FC_AGG_MAX_OUT + o];
      }
      __syncthreads();
      staged = s1 - s0 <= kStage;             // a single pass: the table stays valid for the job's next tiles
      }
      if (active) {
#pragma unroll 1
        for (int k = 0; k < kn; k += kUnroll) {
          float4 l[kUnroll];
#pragma unroll
          for (int u = 0; u < kUnroll; ++u)
            if (k + u < kn) l[u] = ld_stream_f4(s_src[k + u] + idx);
#pragma unroll
          for (int u = 0; u < kUnroll; ++u) {
            if (k + u < kn) {
              const int flag = s_flag[k + u];
              if (flag == FC_AGG_SRC_HOLD) {
                pend = l[u];
              } else {
                float4 x = l[u];
                if (flag == FC_AGG_SRC_MERGE) {   // upload(): W + A*s  (mul, then add; fedavgclient.py:177)
                  const float s = s_scale[k + u];
                  x.x = __fadd_rn(pend.x, __fmul_rn(x.x, s));
                  x.y = __fadd_rn(pend.y, __fmul_rn(x.y, s));
                  x.z = __fadd_rn(pend.z, __fmul_rn(x.z, s));
                  x.w = __fadd_rn(pend.w, __fmul_rn(x.w, s));
                }
#pragma unroll
                for (int o = 0; o < FC_AGG_MAX_OUT; ++o) {
                  const float c = s_coef[k + u][o];
                  if (MODE == FC_AGG_WSUM || c != 0.0f) {   // LERP: c == 0 means "skip" (fedavgserver.py:661)
                    f[o].x = fold1<MODE>(f[o].x, x.x, c);
                    f[o].y = fold1<MODE>(f[o].y, x.y, c);
                    f[o].z = fold1<MODE>(f[o].z, x.z, c);
                    f[o].w = fold1<MODE>(f[o].w, x.w, c);
                  }
                }
              }
            }
          }
        }
      }
    }
    if (active) {
#pragma unroll
      for (int o = 0; o < FC_AGG_MAX_OUT; ++o)
        if (o < nout) st_stream_f4(reinterpret_cast<float*>(P.job_gout[job * FC_AGG_MAX_OUT + o]) + idx, f[o]);
    }
   }
  }
}

}  // namespace

extern "C" int fc_aggregate_tile_floats(void) { return kTileFloats; }

extern "C" int fc_aggregate(int mode, int n_jobs, int n_tiles, const int* job_tile_start,
                            const long long* job_numel, const int* job_nout,
                            const unsigned long long* job_gin, const unsigned long long* job_gout,
                            const float* job_gscale, const int* job_src_start,
                            const unsigned long long* src_ptr, const int* src_flag,
                            const unsigned long long* scale_ptr, const float* coef, int grid_ctas,
                            int device, void* stream) {
  FC_REQUIRE(mode == FC_AGG_LERP || mode == FC_AGG_WSUM, "fc_aggregate: bad mode %d", mode);
  FC_REQUIRE(n_jobs >= 0 && n_tiles >= 0, "fc_aggregate: negative sizes");
  if (n_jobs == 0 || n_tiles == 0) return FC_OK;
  FcDeviceGuard guard(device);
  AggParams P{mode, n_jobs, n_tiles, job_tile_start, job_numel, job_nout, job_gin, job_gout, job_gscale,
              job_src_start, src_ptr, src_flag, scale_ptr, coef};
  int grid = grid_ctas > 0 ? grid_ctas : fc_num_sms(device) * 8;   // 4 resident CTAs/SM x 2 waves
  const int n_runs = (n_tiles + kRun - 1) / kRun;
  if (grid > n_runs) grid = n_runs;
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  if (mode == FC_AGG_LERP)
    aggregate_kernel<FC_AGG_LERP><<<grid, kThreads, 0, st>>>(P);
  else
    aggregate_kernel<FC_AGG_WSUM><<<grid, kThreads, 0, st>>>(P);
  FC_LAUNCH_CHECK();
  return FC_OK;
}
