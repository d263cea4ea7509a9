// tcgen05 / TMEM / TMA GEMM for the client-local transformer (bf16 operands, fp32 accumulate).
//
// Replaces the cuBLAS sgemm calls behind every nn.Linear / CrossModalReparamLinear of the reference's
// Block (/root/reference/src/models/mome.py:58-60,112-121,143-166) and the PatchEmbed conv (:252-265),
// forward and backward, with the elementwise work that follows each of them fused into the epilogue:
//   bias, exact-erf GELU (fwd + derivative), DropPath-scaled residual add, patch-row remap + pos_embed,
//   split-K gradient accumulation.
//
// One CTA = one 128 x 128 output tile (UMMA M=128, N=128, K=16 per instruction), 3-stage TMA->smem ring
// of 64-wide K blocks (128-byte swizzle), accumulator in 128 TMEM columns; 6 warps: TMA producer, MMA
// issuer (+TMEM alloc), 4 epilogue warps (one TMEM lane quarter each).  ~97 KB smem => two CTAs per SM,
// so one CTA's epilogue overlaps the other's main loop (these GEMMs have K = 384..1536: short main
// loops, store-heavy epilogues).
//
// Operand majors: A and B may each be K-major ([rows, K], K contiguous) or MN-major ([K, rows], rows
// contiguous) — the latter serves dW = dY^T X without materialising any transpose.
#include "common.cuh"
#include "sm100.cuh"
#include "../../include/fedcola_b200.h"

#include <mutex>
#include <vector>

namespace {

using namespace sm100;

constexpr int BM = 128, BN = 128, BK = 64, STAGES = 3;
constexpr int A_BYTES = BM * BK * 2, B_BYTES = BN * BK * 2, STAGE_BYTES = A_BYTES + B_BYTES;
constexpr int GEMM_THREADS = 192;
constexpr int SMEM_BYTES = STAGES * STAGE_BYTES + 1024 /*align*/ + 256 /*barriers*/;

struct GemmParams {
  int M, N, K;              // output rows, output cols, reduction length
  int kb_per_split;         // K blocks handled by one blockIdx.z
  int epi;
  int ldo;                  // leading dimension (elements) of out/out2/resid/aux
  void* out;
  void* out2;
  const float* bias;        // [N] or null
  const float* resid;       // EPI_RESID: fp32 [M, ldo]
  const float* row_scale;   // per-group scale (DropPath keep/keep_prob), index = row / rows_per_group; or null
  int rows_per_group;
  const __nv_bfloat16* aux; // EPI_DGELU: pre-activation bf16 [M, ldo]
  const float* pos;         // EPI_PATCH: pos_embed [(P+1), N]
  int patches;              // EPI_PATCH: P (196)
  float alpha;
};

__device__ __forceinline__ float gelu_erf(float x) { return 0.5f * x * (1.0f + erff(x * 0.70710678118654752440f)); }
__device__ __forceinline__ float gelu_erf_grad(float x) {
  const float cdf = 0.5f * (1.0f + erff(x * 0.70710678118654752440f));
  const float pdf = 0.39894228040143267794f * __expf(-0.5f * x * x);
  return cdf + x * pdf;
}
__device__ __forceinline__ uint32_t pack_bf16(float a, float b) {
  __nv_bfloat162 v = __floats2bfloat162_rn(a, b);
  return *reinterpret_cast<uint32_t*>(&v);
}
__device__ __forceinline__ void red_add_v4(float* p, float a, float b, float c, float d) {
  asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(p), "f"(a), "f"(b), "f"(c), "f"(d) : "memory");
}

// One thread owns one output row and 32 consecutive columns [col0, col0+32).
__device__ __forceinline__ void epilogue_chunk(const GemmParams& p, int row, int col0, float (&acc)[32]) {
  if (row >= p.M || col0 >= p.N) return;
  const int ncol = min(32, p.N - col0);        // N is a multiple of 8
  if (p.bias != nullptr) {
#pragma unroll
    for (int i = 0; i < 32; i += 4) {
      if (i < ncol) {
        const float4 b = __ldg(reinterpret_cast<const float4*>(p.bias + col0 + i));
        acc[i] += b.x; acc[i + 1] += b.y; acc[i + 2] += b.z; acc[i + 3] += b.w;
      }
    }
  }
  const size_t o = static_cast<size_t>(row) * p.ldo + col0;
  switch (p.epi) {
    case FC_EPI_BF16: {
      uint4* dst = reinterpret_cast<uint4*>(reinterpret_cast<__nv_bfloat16*>(p.out) + o);
#pragma unroll
      for (int i = 0; i < 32; i += 8)
        if (i < ncol)
          dst[i / 8] = make_uint4(pack_bf16(acc[i], acc[i + 1]), pack_bf16(acc[i + 2], acc[i + 3]),
                                  pack_bf16(acc[i + 4], acc[i + 5]), pack_bf16(acc[i + 6], acc[i + 7]));
    } break;
    case FC_EPI_GELU: {
      uint4* d1 = reinterpret_cast<uint4*>(reinterpret_cast<__nv_bfloat16*>(p.out) + o);
      uint4* d2 = reinterpret_cast<uint4*>(reinterpret_cast<__nv_bfloat16*>(p.out2) + o);
#pragma unroll
      for (int i = 0; i < 32; i += 8)
        if (i < ncol) {
          d1[i / 8] = make_uint4(pack_bf16(acc[i], acc[i + 1]), pack_bf16(acc[i + 2], acc[i + 3]),
                                 pack_bf16(acc[i + 4], acc[i + 5]), pack_bf16(acc[i + 6], acc[i + 7]));
          float g[8];
#pragma unroll
          for (int j = 0; j < 8; ++j) g[j] = gelu_erf(acc[i + j]);
          d2[i / 8] = make_uint4(pack_bf16(g[0], g[1]), pack_bf16(g[2], g[3]), pack_bf16(g[4], g[5]),
                                 pack_bf16(g[6], g[7]));
        }
    } break;
    case FC_EPI_RESID: {
      const float s = p.row_scale ? __ldg(p.row_scale + row / p.rows_per_group) : 1.0f;
      const float4* r = reinterpret_cast<const float4*>(p.resid + o);
      float4* dst = reinterpret_cast<float4*>(reinterpret_cast<float*>(p.out) + o);
#pragma unroll
      for (int i = 0; i < 32; i += 4)
        if (i < ncol) {
          const float4 x = r[i / 4];
          dst[i / 4] = make_float4(x.x + s * acc[i], x.y + s * acc[i + 1], x.z + s * acc[i + 2], x.w + s * acc[i + 3]);
        }
    } break;
    case FC_EPI_DGELU: {
      const uint4* pre = reinterpret_cast<const uint4*>(p.aux + o);
      uint4* dst = reinterpret_cast<uint4*>(reinterpret_cast<__nv_bfloat16*>(p.out) + o);
#pragma unroll
      for (int i = 0; i < 32; i += 8)
        if (i < ncol) {
          const uint4 q = pre[i / 8];
          const __nv_bfloat162* h = reinterpret_cast<const __nv_bfloat162*>(&q);
          float g[8];
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            const float2 x = __bfloat1622float2(h[j]);
            g[2 * j] = acc[i + 2 * j] * gelu_erf_grad(x.x);
            g[2 * j + 1] = acc[i + 2 * j + 1] * gelu_erf_grad(x.y);
          }
          dst[i / 8] = make_uint4(pack_bf16(g[0], g[1]), pack_bf16(g[2], g[3]), pack_bf16(g[4], g[5]),
                                  pack_bf16(g[6], g[7]));
        }
    } break;
    case FC_EPI_F32: {
      float4* dst = reinterpret_cast<float4*>(reinterpret_cast<float*>(p.out) + o);
#pragma unroll
      for (int i = 0; i < 32; i += 4)
        if (i < ncol) dst[i / 4] = make_float4(acc[i], acc[i + 1], acc[i + 2], acc[i + 3]);
    } break;
    case FC_EPI_ATOMIC_F32: {
      float* dst = reinterpret_cast<float*>(p.out) + o;
#pragma unroll
      for (int i = 0; i < 32; i += 4)
        if (i < ncol)
          red_add_v4(dst + i, p.alpha * acc[i], p.alpha * acc[i + 1], p.alpha * acc[i + 2], p.alpha * acc[i + 3]);
    } break;
    case FC_EPI_PATCH: {
      // row = b*P + t  ->  token row b*(P+1) + 1 + t of x; add pos_embed[1+t]
      const int b = row / p.patches, t = row - b * p.patches;
      float4* dst = reinterpret_cast<float4*>(reinterpret_cast<float*>(p.out) +
                                              (static_cast<size_t>(b) * (p.patches + 1) + 1 + t) * p.ldo + col0);
      const float4* pe = reinterpret_cast<const float4*>(p.pos + static_cast<size_t>(1 + t) * p.N + col0);
#pragma unroll
      for (int i = 0; i < 32; i += 4)
        if (i < ncol) {
          const float4 e = __ldg(pe + i / 4);
          dst[i / 4] = make_float4(acc[i] + e.x, acc[i + 1] + e.y, acc[i + 2] + e.z, acc[i + 3] + e.w);
        }
    } break;
    default: break;
  }
}

template <int A_MN, int B_MN>
__global__ void __launch_bounds__(GEMM_THREADS, 2)
gemm_bf16_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB, const GemmParams p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + STAGES * STAGE_BYTES);
  uint64_t* empty_bar = full_bar + STAGES;
  uint64_t* tmem_full_bar = empty_bar + STAGES;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tmem_full_bar + 1);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int n0 = blockIdx.x * BN, m0 = blockIdx.y * BM;
  const int total_kb = (p.K + BK - 1) / BK;
  const int kb0 = blockIdx.z * p.kb_per_split;
  const int kb1 = min(total_kb, kb0 + p.kb_per_split);

  if (warp == 0 && lane == 0) {
    prefetch_tmap(&tmA);
    prefetch_tmap(&tmB);
  }
  if (warp == 1) {
    if (lane == 0) {
      for (int s = 0; s < STAGES; ++s) {
        mbar_init(&full_bar[s], 1);
        mbar_init(&empty_bar[s], 1);
      }
      mbar_init(tmem_full_bar, 1);
      fence_barrier_init();
    }
    __syncwarp();
    tmem_alloc(tmem_slot, BN);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (kb1 > kb0) {
    if (warp == 0) {
      if (lane == 0) {
        int s = 0;
        uint32_t ph = 0;
        for (int kb = kb0; kb < kb1; ++kb) {
          mbar_wait(&empty_bar[s], ph ^ 1);
          mbar_arrive_expect_tx(&full_bar[s], STAGE_BYTES);
          uint8_t* a_dst = smem + s * STAGE_BYTES;
          uint8_t* b_dst = a_dst + A_BYTES;
          if (A_MN) {   // [K, M] global: box {64 (m), 64 (k)} x2
            tma_load_2d(a_dst, &tmA, &full_bar[s], m0, kb * BK);
            tma_load_2d(a_dst + 8192, &tmA, &full_bar[s], m0 + 64, kb * BK);
          } else {      // [M, K] global: box {64 (k), 128 (m)}
            tma_load_2d(a_dst, &tmA, &full_bar[s], kb * BK, m0);
          }
          if (B_MN) {
            tma_load_2d(b_dst, &tmB, &full_bar[s], n0, kb * BK);
            tma_load_2d(b_dst + 8192, &tmB, &full_bar[s], n0 + 64, kb * BK);
          } else {
            tma_load_2d(b_dst, &tmB, &full_bar[s], kb * BK, n0);
          }
          if (++s == STAGES) { s = 0; ph ^= 1; }
        }
      }
    } else if (warp == 1) {
      if (lane == 0) {
        constexpr uint32_t idesc = umma_idesc_bf16(BM, BN, A_MN, B_MN);
        int s = 0;
        uint32_t ph = 0;
        for (int kb = kb0; kb < kb1; ++kb) {
          mbar_wait(&full_bar[s], ph);
          tc_fence_after();
          const uint32_t a_addr = smem_u32(smem + s * STAGE_BYTES);
          const uint32_t b_addr = a_addr + A_BYTES;
#pragma unroll
          for (int k = 0; k < BK / 16; ++k) {
            // K-major: 16 K-elements = 32 B inside the 128-B swizzle row; SBO = 8 rows * 128 B.
            // MN-major: 16 K-rows of 128 B = 2048 B; LBO = next 64-wide MN block (8 KB), SBO = 8 K-rows.
            const uint64_t ad = A_MN ? umma_smem_desc(a_addr + k * 2048, 8192, 1024)
                                     : umma_smem_desc(a_addr + k * 32, 16, 1024);
            const uint64_t bd = B_MN ? umma_smem_desc(b_addr + k * 2048, 8192, 1024)
                                     : umma_smem_desc(b_addr + k * 32, 16, 1024);
            umma_bf16(tmem_base, ad, bd, idesc, (kb > kb0 || k > 0) ? 1u : 0u);
          }
          umma_commit(&empty_bar[s]);          // smem slot free once these MMAs retire
          if (++s == STAGES) { s = 0; ph ^= 1; }
        }
        umma_commit(tmem_full_bar);            // accumulator complete
      }
    } else {
      const int q = warp & 3;                  // TMEM lane quarter this warp may access
      mbar_wait(tmem_full_bar, 0);
      tc_fence_after();
      const int row = m0 + q * 32 + lane;
#pragma unroll 1
      for (int c = 0; c < BN / 32; ++c) {
        float acc[32];
        tmem_ld_32x32(tmem_base + (static_cast<uint32_t>(q * 32) << 16) + c * 32, acc);
        tmem_ld_wait();
        epilogue_chunk(p, row, n0 + c * 32, acc);
      }
      tc_fence_before();
    }
  }
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, BN);
  }
}

// ---- host side ------------------------------------------------------------------------------------
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

EncodeTiledFn get_encode_fn() {
  static EncodeTiledFn fn = nullptr;
  static std::once_flag once;
  std::call_once(once, [] {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess &&
        q == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<EncodeTiledFn>(p);
  });
  return fn;
}

// bf16 row-major matrix [rows, cols] with leading dimension ld; box = {box_cols, box_rows}
int make_tmap(CUtensorMap* m, const void* ptr, long long rows, long long cols, long long ld, int box_cols,
              int box_rows) {
  EncodeTiledFn fn = get_encode_fn();
  if (!fn) FC_FAIL(FC_ERR_CUDA, "cuTensorMapEncodeTiled entry point unavailable");
  cuuint64_t gdim[2] = {static_cast<cuuint64_t>(cols), static_cast<cuuint64_t>(rows)};
  cuuint64_t gstr[1] = {static_cast<cuuint64_t>(ld) * 2};
  cuuint32_t box[2] = {static_cast<cuuint32_t>(box_cols), static_cast<cuuint32_t>(box_rows)};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = fn(m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(ptr), gdim, gstr, box, estr,
                  CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                  CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) FC_FAIL(FC_ERR_CUDA, "cuTensorMapEncodeTiled failed (%d) rows=%lld cols=%lld ld=%lld", (int)r, rows, cols, ld);
  return FC_OK;
}

// Optional per-launch timing (bench.py's roofline): CUDA events on the launching stream around every GEMM.
struct ProfRec { cudaEvent_t a, b; double flops; };
std::mutex g_prof_mu;
std::vector<ProfRec> g_prof;
int g_prof_on = 0;

template <int A_MN, int B_MN>
int launch(const CUtensorMap& ta, const CUtensorMap& tb, const GemmParams& p, int splits, cudaStream_t st) {
  auto kern = gemm_bf16_kernel<A_MN, B_MN>;
  // the >48 KB dynamic-smem opt-in is per device; remember which device this thread last configured
  static thread_local int configured_dev = -1;
  int dev = -1;
  cudaGetDevice(&dev);
  if (dev != configured_dev) {
    FC_CUDA_CHECK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_BYTES));
    configured_dev = dev;
  }
  dim3 grid((p.N + BN - 1) / BN, (p.M + BM - 1) / BM, splits);
  ProfRec rec{nullptr, nullptr, 2.0 * p.M * (double)p.N * p.K};
  const bool prof = __atomic_load_n(&g_prof_on, __ATOMIC_RELAXED) != 0;
  if (prof) {
    cudaEventCreate(&rec.a);
    cudaEventCreate(&rec.b);
    cudaEventRecord(rec.a, st);
  }
  kern<<<grid, GEMM_THREADS, SMEM_BYTES, st>>>(ta, tb, p);
  if (prof) {
    cudaEventRecord(rec.b, st);
    std::lock_guard<std::mutex> lk(g_prof_mu);
    g_prof.push_back(rec);
  }
  FC_LAUNCH_CHECK();
  return FC_OK;
}

}  // namespace

extern "C" void fc_gemm_profile(int enable) {
  std::lock_guard<std::mutex> lk(g_prof_mu);
  for (auto& r : g_prof) { cudaEventDestroy(r.a); cudaEventDestroy(r.b); }
  g_prof.clear();
  __atomic_store_n(&g_prof_on, enable, __ATOMIC_RELAXED);
}

// Sums the recorded launches (synchronises on their events). Returns the number of launches.
extern "C" long long fc_gemm_profile_collect(double* total_ms, double* total_flops) {
  std::lock_guard<std::mutex> lk(g_prof_mu);
  double ms = 0.0, fl = 0.0;
  for (auto& r : g_prof) {
    cudaEventSynchronize(r.b);
    float t = 0.f;
    if (cudaEventElapsedTime(&t, r.a, r.b) == cudaSuccess) ms += t;
    fl += r.flops;
  }
  if (total_ms) *total_ms = ms;
  if (total_flops) *total_flops = fl;
  return (long long)g_prof.size();
}

extern "C" int fc_gemm_bf16(int M, int N, int K, const void* A, long long lda, int a_mn_major, const void* B,
                            long long ldb, int b_mn_major, int epi, void* out, void* out2, long long ldo,
                            const float* bias, const float* resid, const float* row_scale, int rows_per_group,
                            const void* aux, const float* pos, int patches, float alpha, int splits, int device,
                            void* stream) {
  FC_REQUIRE(M > 0 && N > 0 && K > 0, "fc_gemm_bf16: empty problem %d %d %d", M, N, K);
  FC_REQUIRE(N % 8 == 0 && lda % 8 == 0 && ldb % 8 == 0 && ldo % 4 == 0, "fc_gemm_bf16: N, lda, ldb must be multiples of 8");
  FC_REQUIRE(epi >= FC_EPI_BF16 && epi <= FC_EPI_PATCH, "fc_gemm_bf16: bad epilogue %d", epi);
  FC_REQUIRE(out != nullptr, "fc_gemm_bf16: null output");
  FC_REQUIRE(epi != FC_EPI_GELU || out2 != nullptr, "fc_gemm_bf16: GELU epilogue needs out2");
  FC_REQUIRE(epi != FC_EPI_RESID || resid != nullptr, "fc_gemm_bf16: RESID epilogue needs resid");
  FC_REQUIRE(epi != FC_EPI_DGELU || aux != nullptr, "fc_gemm_bf16: DGELU epilogue needs aux");
  FC_REQUIRE(epi != FC_EPI_PATCH || (pos != nullptr && patches > 0), "fc_gemm_bf16: PATCH epilogue needs pos");
  FC_REQUIRE(row_scale == nullptr || rows_per_group > 0, "fc_gemm_bf16: rows_per_group");
  FcDeviceGuard guard(device);
  const int total_kb = (K + BK - 1) / BK;
  if (splits < 1) splits = 1;
  if (splits > total_kb) splits = total_kb;
  FC_REQUIRE(splits == 1 || epi == FC_EPI_ATOMIC_F32, "fc_gemm_bf16: split-K needs the atomic epilogue");
  GemmParams p;
  p.M = M; p.N = N; p.K = K;
  p.kb_per_split = (total_kb + splits - 1) / splits;
  splits = (total_kb + p.kb_per_split - 1) / p.kb_per_split;
  p.epi = epi; p.ldo = static_cast<int>(ldo);
  p.out = out; p.out2 = out2; p.bias = bias; p.resid = resid; p.row_scale = row_scale;
  p.rows_per_group = rows_per_group > 0 ? rows_per_group : 1;
  p.aux = reinterpret_cast<const __nv_bfloat16*>(aux); p.pos = pos; p.patches = patches; p.alpha = alpha;
  CUtensorMap ta, tb;
  int rc;
  // K-major operand: global [rows, K]; MN-major operand: global [K, rows].
  rc = a_mn_major ? make_tmap(&ta, A, K, M, lda, 64, 64) : make_tmap(&ta, A, M, K, lda, 64, BM);
  if (rc) return rc;
  rc = b_mn_major ? make_tmap(&tb, B, K, N, ldb, 64, 64) : make_tmap(&tb, B, N, K, ldb, 64, BN);
  if (rc) return rc;
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  if (a_mn_major && b_mn_major) return launch<1, 1>(ta, tb, p, splits, st);
  if (a_mn_major) return launch<1, 0>(ta, tb, p, splits, st);
  if (b_mn_major) return launch<0, 1>(ta, tb, p, splits, st);
  return launch<0, 0>(ta, tb, p, splits, st);
}
