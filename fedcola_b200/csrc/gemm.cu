// tcgen05 / TMEM / TMA GEMM for the client-local transformer (bf16 operands, fp32 accumulate).
//
// Replaces the cuBLAS sgemm calls behind every nn.Linear / CrossModalReparamLinear of the reference's
// Block (/root/reference/src/models/mome.py:58-60,112-121,143-166) and the PatchEmbed conv (:252-265),
// forward and backward, with the elementwise work that follows each of them fused into the epilogue:
//   bias, exact-erf GELU (value + derivative in one pass), multiply-by-saved-derivative (+ bias-gradient
//   column sums), DropPath-scaled residual add, patch-row remap + pos_embed, split-K gradient accumulation.
//
// Persistent, warp-specialised kernel: one CTA per SM walks 128 x BN output tiles (BN = 128 / 192 / 256,
// UMMA M=128, N=BN, K=16 per instruction).  A 3-5 stage TMA->smem ring of 64-wide K blocks (128-byte
// swizzle) runs across tile boundaries; the fp32 accumulator is double-buffered in TMEM (2 x 256 columns)
// so the 16 epilogue warps drain tile i while the MMA warp already works on tile i+1 — these GEMMs have
// K = 384..1536: short main loops, store-heavy epilogues.  The epilogue (the measured bottleneck: it is
// instruction-issue bound) is specialised per fused op at compile time, transposes each 32x32 accumulator
// block through shared memory (two 16-row passes) so that every global load/store of a warp covers whole
// 64/128-byte row segments, and issues the global loads it needs (bias, residual, saved gelu') before the
// TMEM load so that their latency overlaps it.
//
// Operand majors: A and B may each be K-major ([rows, K], K contiguous) or MN-major ([K, rows], rows
// contiguous) — the latter serves dX = dY W and dW = dY^T X without materialising any transpose.
#include "common.cuh"
#include "sm100.cuh"
#include "../../include/fedcola_b200.h"

#include <cstdlib>
#include <cstring>
#include <mutex>
#include <vector>

namespace {

using namespace sm100;

constexpr int BM = 128, BK = 64;
constexpr int A_BYTES = BM * BK * 2;
// Epilogue warps (per TMEM lane quarter: n/4 warps, interleaved column chunks).  The shared memory their staging
// tiles do not take goes to the TMA ring, and at K = 384 the main loop is limited by the bytes the ring keeps in
// flight per SM: epilogues with little arithmetic run 8 warps + one more stage (main loop 10-16 % faster), the
// GELU / GELU'-multiply / patch epilogues need all 16 (measured with tools/gemm_bench.py).
constexpr int epi_warps_for(int epi) {
  return (epi == FC_EPI_GELU || epi == FC_EPI_MULAUX || epi == FC_EPI_PATCH) ? 16 : 8;
}
constexpr int gemm_threads_for(int epi) { return 64 + epi_warps_for(epi) * 32; }   // + TMA producer, MMA issuer
constexpr int STAGE_LD = 36;               // floats per row of the legacy transpose buffer (PATCH epilogue only)
constexpr int EPI_BUF_BYTES = 32 * 128;    // per epilogue warp: 32 rows x 128 B, 128B-swizzled (TMA store/load tile)
constexpr int EPI_BIAS_BYTES = 64 * 4;      // per epilogue warp: the bias of the chunk's (up to) 64 columns
constexpr int TMEM_BUF_COLS = 256;         // two accumulator buffers at columns 0 and 256

template <int BN, int EPI, int PAIR> struct Cfg {
  static constexpr int EPI_WARPS = epi_warps_for(EPI);
  static constexpr int EPI_STAGE_BYTES = EPI_WARPS * (EPI_BUF_BYTES + EPI_BIAS_BYTES);
  static constexpr int B_BYTES = BN * BK * 2;
  // pair mode (cta_group::2): a stage holds this CTA's 128 rows of A and HALF of the B tile (BN/2 rows / columns)
  static constexpr int STAGE_BYTES = A_BYTES + (PAIR ? B_BYTES / 2 : B_BYTES);
  // as many ring stages as fit beside the epilogue staging (1 KB of barriers / slack), at most 8
  static constexpr int FIT = (232448 - EPI_STAGE_BYTES - 1024) / STAGE_BYTES;
  static constexpr int STAGES = FIT > 8 ? 8 : FIT;
  static constexpr int SMEM_BYTES = STAGES * STAGE_BYTES + EPI_STAGE_BYTES + 1024 /*barriers*/;
  static_assert(STAGES >= 3, "too few pipeline stages");
  static_assert(SMEM_BYTES <= 232448, "exceeds the 227 KB of shared memory a CTA may use");
  static_assert((2 * STAGES + 4 + EPI_WARPS) * 8 + 16 <= 1024, "barrier block");
};

// Client groups: one launch can run the SAME GEMM (shape, majors, epilogue) for up to FC_GEMM_MAX_GROUPS independent
// operand sets — the same layer of several clients trained in lockstep.  Tiles of all groups share one persistent
// walk, so the per-launch fixed costs (prologue, first-load latency, exposed last epilogue, wave quantisation) are
// paid once per group of clients instead of once per client.
struct GroupPtrs {
  void* out;
  void* out2;
  const float* bias;        // [N] or null
  const float* resid;       // EPI_RESID: fp32 [M, ldo]
  const float* row_scale;   // per-sample scale (DropPath keep/keep_prob), index = row / rows_per_group; or null
  const __nv_bfloat16* aux; // EPI_MULAUX: bf16 multiplier [M, ldo] (gelu'(pre) saved by the forward)
  float* colsum;            // EPI_MULAUX: += column sums of the output (bias gradient), or null
  const float* pos;         // EPI_PATCH: pos_embed [(P+1), N]
};
struct GroupMaps { CUtensorMap a, b, o, o2, r, a_lo, b_lo; };   // a_lo / b_lo: split-operand (fp32-accurate) mode
struct AllMaps { GroupMaps g[FC_GEMM_MAX_GROUPS]; };

struct GemmParams {
  int M, N, K;              // output rows, output cols, reduction length (of every group)
  int kb_per_split, splits; // K blocks handled by one split; number of splits
  int m_tiles, n_tiles;
  int groups;
  int epi;
  int ldo;                  // leading dimension (elements) of out/out2/resid/aux
  int rows_per_group;
  int patches;              // EPI_PATCH: P (196)
  float alpha;
  int debug;                // measurement aid (FC_GEMM_DEBUG env): 1 = skip epilogue work, 2 = skip TMA+MMA work
  int pair;                 // 1: CTA pairs (cta_group::2) compute 256 x BN tiles, each CTA stages its 128 rows of A and half of B
  int split;                // 1: split operands A = A_hi + A_lo, B = B_hi + B_lo (bf16 pairs): acc = A_hi B_hi + A_hi B_lo + A_lo B_hi
  GroupPtrs g[FC_GEMM_MAX_GROUPS];
};

// Persistent tile walk.  pair == 0: tile t = blockIdx.x, +gridDim.x, ... -> (split, m_blk, n_blk), n fastest.
// pair == 1: the CTA pair (blockIdx.x >> 1) walks pair-tiles (split, m_pair, n_blk); rank r owns m_blk = 2*m_pair + r
// (an M tile past the matrix is all TMA zero fill / clipped stores).
template <int PAIR> struct TileWalk {
  int first, step, total, tiles_mn, per_group, n_tiles, rank;
  __device__ TileWalk(const GemmParams& p) {
    rank = PAIR ? static_cast<int>(blockIdx.x & 1) : 0;
    first = PAIR ? static_cast<int>(blockIdx.x >> 1) : static_cast<int>(blockIdx.x);
    step = PAIR ? static_cast<int>(gridDim.x >> 1) : static_cast<int>(gridDim.x);
    n_tiles = p.n_tiles;
    tiles_mn = (PAIR ? (p.m_tiles + 1) / 2 : p.m_tiles) * p.n_tiles;
    per_group = tiles_mn * p.splits;
    total = per_group * p.groups;
  }
  // tile -> (group, split, m_blk, n_blk): n fastest (CTAs running side by side share the A rows in L2), group slowest
  __device__ void decode(int tile, int& grp, int& split, int& m_blk, int& n_blk) const {
    grp = tile / per_group;
    const int t = tile - grp * per_group;
    split = t / tiles_mn;
    const int mn = t - split * tiles_mn;
    const int mq = mn / n_tiles;
    n_blk = mn - mq * n_tiles;
    m_blk = PAIR ? 2 * mq + rank : mq;
  }
};

// Exact-erf GELU (nn.GELU() default) evaluated with the Abramowitz-Stegun 7.1.26 rational form of erf
// (|error| <= 1.5e-7, far below the bf16 rounding of the stored result): one MUFU.RCP + one MUFU.EX2 + 6 FMAs
// instead of erff's ~30-instruction polynomial.
//   Phi(x) = 0.5*(1 + erf(x/sqrt2));  with z = |x|/sqrt2, t = 1/(1 + p z):  1 - erf(z) = poly(t) * exp(-z^2)
__device__ __forceinline__ void gelu_parts(float x, float& cdf, float& e) {
  const float z = fabsf(x) * 0.70710678118654752440f;
  float t;
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(t) : "f"(fmaf(0.3275911f, z, 1.0f)));      // MUFU.RCP (2 ulp)
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e) : "f"(-1.4426950408889634f * z * z));      // exp(-z^2) = exp(-x^2/2)
  float poly = fmaf(t, 1.061405429f, -1.453152027f);
  poly = fmaf(t, poly, 1.421413741f);
  poly = fmaf(t, poly, -0.284496736f);
  poly = fmaf(t, poly, 0.254829592f);
  const float tail = 0.5f * t * poly * e;                         // = 0.5*(1 - erf(z)) = Phi(-|x|), no cancellation
  cdf = x >= 0.0f ? 1.0f - tail : tail;
}
// GELU of two adjacent columns: g = x*Phi(x), d = gelu'(x) = Phi(x) + x*phi(x).  Same Abramowitz-Stegun form as
// gelu_parts (coefficients pre-multiplied by 0.5; Phi(x) = 0.5 + copysign(0.5 - Phi(-|x|), x)).
__device__ __forceinline__ void gelu_pair(float x0, float x1, uint64_t& g, uint64_t& d) {
  const uint64_t x = f2_pack(x0, x1);
  const uint64_t den = f2_fma(f2_pack(fabsf(x0), fabsf(x1)), f2_all(0.3275911f * 0.70710678118654752440f), f2_all(1.0f));
  const uint64_t arg = f2_mul(f2_mul(x, x), f2_all(-0.5f * 1.4426950408889634f));
  float d0, d1, a0, a1, t0, t1, e0, e1;
  f2_unpack(den, d0, d1);
  f2_unpack(arg, a0, a1);
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(t0) : "f"(d0));
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(t1) : "f"(d1));
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e0) : "f"(a0));
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e1) : "f"(a1));
  const uint64_t t = f2_pack(t0, t1), e = f2_pack(e0, e1);
  uint64_t poly = f2_fma(t, f2_all(0.5f * 1.061405429f), f2_all(0.5f * -1.453152027f));
  poly = f2_fma(t, poly, f2_all(0.5f * 1.421413741f));
  poly = f2_fma(t, poly, f2_all(0.5f * -0.284496736f));
  poly = f2_fma(t, poly, f2_all(0.5f * 0.254829592f));
  const uint64_t tail = f2_mul(f2_mul(t, poly), e);                 // Phi(-|x|) in [0, 0.5]
  const uint64_t half_m = f2_fma(tail, f2_all(-1.0f), f2_all(0.5f)); // 0.5 - Phi(-|x|) >= 0
  float h0, h1;
  f2_unpack(half_m, h0, h1);
  h0 = __uint_as_float(__float_as_uint(h0) | (__float_as_uint(x0) & 0x80000000u));
  h1 = __uint_as_float(__float_as_uint(h1) | (__float_as_uint(x1) & 0x80000000u));
  const uint64_t cdf = f2_add(f2_pack(h0, h1), f2_all(0.5f));
  g = f2_mul(x, cdf);
  d = f2_fma(f2_mul(x, e), f2_all(0.39894228040143267794f), cdf);
}
__device__ __forceinline__ uint32_t pack_bf16_f2(uint64_t v) {
  float a, b;
  f2_unpack(v, a, b);
  __nv_bfloat162 r = __floats2bfloat162_rn(a, b);
  return *reinterpret_cast<uint32_t*>(&r);
}
__device__ __forceinline__ uint32_t pack_bf16(float a, float b) {
  __nv_bfloat162 v = __floats2bfloat162_rn(a, b);
  return *reinterpret_cast<uint32_t*>(&v);
}
// explicit shared-state-space accesses (32-bit addresses): the transpose buffer must not go through generic LD/ST
__device__ __forceinline__ void sts_v4(uint32_t addr, float a, float b, float c, float d) {
  asm volatile("st.shared.v4.f32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "f"(a), "f"(b), "f"(c), "f"(d) : "memory");
}
__device__ __forceinline__ float4 lds_v4(uint32_t addr) {
  float4 v;
  asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(addr) : "memory");
  return v;
}
__device__ __forceinline__ void red_add_v4(float* p, float a, float b, float c, float d) {
  asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(p), "f"(a), "f"(b), "f"(c), "f"(d) : "memory");
}

// Operands the epilogue reads from global memory (residual rows / saved gelu'), fetched for a whole 32x32
// chunk (both 16-row passes) BEFORE the TMEM load + transpose so that their latency overlaps that work.
// Entry 4*h + j belongs to row  row_base + (lane>>3) + 16*h + 4*j  (h = pass, j = 0..3).
template <int EPI> struct EpiPre {};
template <> struct EpiPre<FC_EPI_RESID> { float4 x[8]; float sc[8]; };
template <> struct EpiPre<FC_EPI_MULAUX> { uint2 q[8]; };   // (legacy register-prefetch path; PATCH is its only user now)

template <int EPI>
__device__ __forceinline__ void epilogue_prefetch(const GemmParams& p, const GroupPtrs& gp, EpiPre<EPI>& pre, int row_base, int col0, int lane) {
  const int col = col0 + (lane & 7) * 4;
  const int r0 = row_base + (lane >> 3);
  if constexpr (EPI == FC_EPI_RESID) {
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      const int row = r0 + 4 * i;
      pre.x[i] = make_float4(0.f, 0.f, 0.f, 0.f);
      pre.sc[i] = 1.0f;
      if (col < p.N && row < p.M) {
        pre.x[i] = *reinterpret_cast<const float4*>(gp.resid + static_cast<size_t>(row) * p.ldo + col);
        if (gp.row_scale) pre.sc[i] = __ldg(gp.row_scale + row / p.rows_per_group);
      }
    }
  } else if constexpr (EPI == FC_EPI_MULAUX) {
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      const int row = r0 + 4 * i;
      pre.q[i] = make_uint2(0u, 0u);
      if (col < p.N && row < p.M) pre.q[i] = *reinterpret_cast<const uint2*>(gp.aux + static_cast<size_t>(row) * p.ldo + col);
    }
  }
}

// Epilogue of one 16x32 accumulator block that sits transposed in `stage` (row-major, STAGE_LD floats per
// row).  Lane l owns columns col..col+3 (col = col0 + 4*(l&7)) of rows r_i = (l>>3) + 4*i, i = 0..3: every
// global access of the warp covers 4 rows x (128 B fp32 | 64 B bf16) contiguous segments.
// EPI is a compile-time constant: one lean instruction stream per fused op.  H = which 16-row pass.
template <int EPI, int H>
__device__ __forceinline__ void epilogue_block(const GemmParams& p, const GroupPtrs& gp, uint32_t stage, int row_base, int col0, int lane,
                                               float4 bias, const EpiPre<EPI>& pre) {
  const int rsub = lane >> 3, col = col0 + (lane & 7) * 4;
  const bool col_ok = col < p.N;                 // N is a multiple of 8, col a multiple of 4
  if (EPI != FC_EPI_MULAUX && !col_ok) return;   // (MULAUX keeps the whole warp for the colsum shuffles)
  float4 v[4];
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    v[i] = lds_v4(stage + ((rsub + 4 * i) * STAGE_LD + (lane & 7) * 4) * 4);
    v[i].x += bias.x; v[i].y += bias.y; v[i].z += bias.z; v[i].w += bias.w;
  }
  const int r0 = row_base + H * 16 + rsub;
  bool ok[4];
#pragma unroll
  for (int i = 0; i < 4; ++i) ok[i] = col_ok && (r0 + 4 * i < p.M);
  const size_t o0 = static_cast<size_t>(r0) * p.ldo + col;
  const size_t rstep = static_cast<size_t>(4) * p.ldo;

  if constexpr (EPI == FC_EPI_BF16) {
    __nv_bfloat16* out = reinterpret_cast<__nv_bfloat16*>(gp.out) + o0;
#pragma unroll
    for (int i = 0; i < 4; ++i)
      if (ok[i]) *reinterpret_cast<uint2*>(out + i * rstep) = make_uint2(pack_bf16(v[i].x, v[i].y), pack_bf16(v[i].z, v[i].w));
  } else if constexpr (EPI == FC_EPI_GELU) {
    // out = gelu'(pre) (what the backward needs), out2 = gelu(pre) (the fc2 operand); Phi and exp are shared
    __nv_bfloat16* o1 = reinterpret_cast<__nv_bfloat16*>(gp.out) + o0;
    __nv_bfloat16* o2 = reinterpret_cast<__nv_bfloat16*>(gp.out2) + o0;
#pragma unroll
    for (int i = 0; i < 4; ++i)
      if (ok[i]) {
        float g[4], d[4];
        const float x[4] = {v[i].x, v[i].y, v[i].z, v[i].w};
#pragma unroll
        for (int k = 0; k < 4; ++k) {
          float cdf, e;
          gelu_parts(x[k], cdf, e);
          g[k] = x[k] * cdf;
          d[k] = fmaf(x[k] * 0.39894228040143267794f, e, cdf);     // Phi(x) + x*phi(x)
        }
        *reinterpret_cast<uint2*>(o1 + i * rstep) = make_uint2(pack_bf16(d[0], d[1]), pack_bf16(d[2], d[3]));
        *reinterpret_cast<uint2*>(o2 + i * rstep) = make_uint2(pack_bf16(g[0], g[1]), pack_bf16(g[2], g[3]));
      }
  } else if constexpr (EPI == FC_EPI_RESID) {
    float* out = reinterpret_cast<float*>(gp.out) + o0;
#pragma unroll
    for (int i = 0; i < 4; ++i)
      if (ok[i]) {
        const float4 x = pre.x[H * 4 + i];
        const float sc = pre.sc[H * 4 + i];
        *reinterpret_cast<float4*>(out + i * rstep) =
            make_float4(x.x + sc * v[i].x, x.y + sc * v[i].y, x.z + sc * v[i].z, x.w + sc * v[i].w);
      }
  } else if constexpr (EPI == FC_EPI_MULAUX) {
    __nv_bfloat16* out = reinterpret_cast<__nv_bfloat16*>(gp.out) + o0;
    float4 cs = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
    for (int i = 0; i < 4; ++i)
      if (ok[i]) {
        const uint2 q = pre.q[H * 4 + i];
        const float2 lo = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&q.x));
        const float2 hi = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&q.y));
        const uint32_t w0 = pack_bf16(v[i].x * lo.x, v[i].y * lo.y), w1 = pack_bf16(v[i].z * hi.x, v[i].w * hi.y);
        *reinterpret_cast<uint2*>(out + i * rstep) = make_uint2(w0, w1);
        // sum what was actually stored (bf16-rounded), as a separate column-sum pass over the output would
        const float2 a = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&w0));
        const float2 b = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&w1));
        cs.x += a.x; cs.y += a.y; cs.z += b.x; cs.w += b.y;
      }
    if (gp.colsum != nullptr) {            // 16 rows of this block: lanes l, l+8, l+16, l+24 share the columns
#pragma unroll
      for (int o = 8; o <= 16; o <<= 1) {
        cs.x += __shfl_xor_sync(0xffffffffu, cs.x, o); cs.y += __shfl_xor_sync(0xffffffffu, cs.y, o);
        cs.z += __shfl_xor_sync(0xffffffffu, cs.z, o); cs.w += __shfl_xor_sync(0xffffffffu, cs.w, o);
      }
      if (lane < 8 && col_ok) red_add_v4(gp.colsum + col, cs.x, cs.y, cs.z, cs.w);
    }
  } else if constexpr (EPI == FC_EPI_F32) {
    float* out = reinterpret_cast<float*>(gp.out) + o0;
#pragma unroll
    for (int i = 0; i < 4; ++i)
      if (ok[i]) *reinterpret_cast<float4*>(out + i * rstep) = v[i];
  } else if constexpr (EPI == FC_EPI_ATOMIC_F32) {
    float* out = reinterpret_cast<float*>(gp.out) + o0;
#pragma unroll
    for (int i = 0; i < 4; ++i)
      if (ok[i]) red_add_v4(out + i * rstep, p.alpha * v[i].x, p.alpha * v[i].y, p.alpha * v[i].z, p.alpha * v[i].w);
  } else if constexpr (EPI == FC_EPI_PATCH) {
    // row = b*P + t  ->  token row b*(P+1) + 1 + t of x; add pos_embed[1+t]
    float4 e[4];
    int orow[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      e[i] = make_float4(0.f, 0.f, 0.f, 0.f);
      orow[i] = 0;
      if (ok[i]) {
        const int row = r0 + 4 * i, b = row / p.patches, t = row - b * p.patches;
        orow[i] = b * (p.patches + 1) + 1 + t;
        e[i] = __ldg(reinterpret_cast<const float4*>(gp.pos + static_cast<size_t>(1 + t) * p.N + col));
      }
    }
    float* out = reinterpret_cast<float*>(gp.out);
#pragma unroll
    for (int i = 0; i < 4; ++i)
      if (ok[i])
        *reinterpret_cast<float4*>(out + static_cast<size_t>(orow[i]) * p.ldo + col) =
            make_float4(v[i].x + e[i].x, v[i].y + e[i].y, v[i].z + e[i].z, v[i].w + e[i].w);
  }
}

// ---- TMA-driven epilogue ------------------------------------------------------------------------------
// Each epilogue warp owns a 32-row x 128-byte staging tile in shared memory, laid out exactly as a
// SWIZZLE_128B TMA box: 16-byte chunk j of row t sits at  t*128 + ((j ^ (t & 7)) << 4).  A thread owns one
// accumulator row (TMEM lane), so it writes its row straight into the tile (conflict-free thanks to the
// swizzle); ONE thread then hands the tile to the TMA engine (cp.async.bulk.tensor store / reduce-add), which
// does address generation, coalescing and M/N boundary clipping.  Residual / saved-gelu' operands arrive the same
// way (TMA load into the tile, combined in place).  ~0.06 instructions per output element instead of ~0.4.
__device__ __forceinline__ uint32_t swz128(uint32_t buf, int row, int chunk) {
  return buf + row * 128 + ((chunk ^ (row & 7)) << 4);
}
__device__ __forceinline__ void sts_u4(uint32_t addr, uint32_t a, uint32_t b, uint32_t c, uint32_t d) {
  asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "r"(a), "r"(b), "r"(c), "r"(d) : "memory");
}
__device__ __forceinline__ uint4 lds_u4(uint32_t addr) {
  uint4 v;
  asm volatile("ld.shared.v4.b32 {%0, %1, %2, %3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "r"(addr) : "memory");
  return v;
}
__device__ __forceinline__ uint32_t lds_u32(uint32_t addr) {
  uint32_t v;
  asm volatile("ld.shared.b32 %0, [%1];" : "=r"(v) : "r"(addr) : "memory");
  return v;
}
__device__ __forceinline__ float2 bf2_to_f2(uint32_t w) {
  return __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&w));
}

template <int EPI> struct EpiTraits {
  static constexpr bool kOut16 = (EPI == FC_EPI_BF16 || EPI == FC_EPI_GELU || EPI == FC_EPI_MULAUX);
  static constexpr int kCW = kOut16 ? 64 : 32;      // columns per staging tile (128-byte rows either way)
  static constexpr bool kLoads = (EPI == FC_EPI_RESID || EPI == FC_EPI_MULAUX);
};

// Store the staged tile: generic-proxy writes -> async proxy, one elected lane issues.  The store drains in the
// background; stage_acquire() must be called before the staging buffer is written (or TMA-loaded) again.
template <bool REDUCE>
__device__ __forceinline__ void stage_store(const CUtensorMap* map, const void* buf_ptr, int col0, int row0, int lane) {
  fence_proxy_async();
  __syncwarp();
  // elect.sync picks the same lane for the same (full) mask every time: the bulk-group wait in stage_acquire() is
  // executed by the thread that committed the store.  (`lane == 0` would put the UTMASTG in a BRA.U.ANY loop.)
  if (elect_one()) {
    if (REDUCE) tma_reduce_add_2d(map, buf_ptr, col0, row0); else tma_store_2d(map, buf_ptr, col0, row0);
    tma_store_commit();
  }
}
__device__ __forceinline__ void stage_acquire(int lane) {
  if (elect_one()) tma_store_wait_read();   // previous store of this warp has finished reading the buffer
  __syncwarp();
}

// One 32-row x kCW-column chunk of the accumulator (acc[] = this lane's row) -> global, fused op EPI.
template <int EPI>
__device__ __forceinline__ void epilogue_tma_chunk(const GemmParams& p, const GroupPtrs& gp, const CUtensorMap* tmO, const CUtensorMap* tmO2,
                                                   uint8_t* buf_ptr, uint32_t buf, uint32_t bias_smem, uint64_t* ebar,
                                                   uint32_t& eph, float (&acc)[EpiTraits<EPI>::kCW], int row0, int col0,
                                                   int lane) {
  constexpr int CW = EpiTraits<EPI>::kCW;
  const int t = lane;
  // bias of the chunk's columns was staged in shared memory by this warp (one coalesced load, issued before the
  // TMEM read): every lane needs all of it -> broadcast LDS.128
  if (gp.bias != nullptr) {
#pragma unroll
    for (int j = 0; j < CW / 4; ++j) {
      const float4 b = lds_v4(bias_smem + 16 * j);
      acc[4 * j] += b.x; acc[4 * j + 1] += b.y; acc[4 * j + 2] += b.z; acc[4 * j + 3] += b.w;
    }
  }
  if constexpr (EPI == FC_EPI_BF16) {
    stage_acquire(lane);
#pragma unroll
    for (int j = 0; j < 8; ++j)
      sts_u4(swz128(buf, t, j), pack_bf16(acc[8 * j], acc[8 * j + 1]), pack_bf16(acc[8 * j + 2], acc[8 * j + 3]),
             pack_bf16(acc[8 * j + 4], acc[8 * j + 5]), pack_bf16(acc[8 * j + 6], acc[8 * j + 7]));
    stage_store<false>(tmO, buf_ptr, col0, row0, lane);
  } else if constexpr (EPI == FC_EPI_GELU) {
    // out = gelu'(x) (saved for the backward), out2 = gelu(x) (the fc2 operand); Phi and exp are shared
    uint32_t gp[32];
    stage_acquire(lane);
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      uint32_t dp[4];
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        uint64_t g2, d2;
        gelu_pair(acc[8 * j + 2 * k], acc[8 * j + 2 * k + 1], g2, d2);
        gp[4 * j + k] = pack_bf16_f2(g2);
        dp[k] = pack_bf16_f2(d2);
      }
      sts_u4(swz128(buf, t, j), dp[0], dp[1], dp[2], dp[3]);
    }
    stage_store<false>(tmO, buf_ptr, col0, row0, lane);
    stage_acquire(lane);
#pragma unroll
    for (int j = 0; j < 8; ++j) sts_u4(swz128(buf, t, j), gp[4 * j], gp[4 * j + 1], gp[4 * j + 2], gp[4 * j + 3]);
    stage_store<false>(tmO2, buf_ptr, col0, row0, lane);
  } else if constexpr (EPI == FC_EPI_RESID) {
    const int row = row0 + t;
    const float sc = (gp.row_scale != nullptr && row < p.M) ? __ldg(gp.row_scale + row / p.rows_per_group) : 1.0f;
    mbar_wait(ebar, eph);                       // residual tile has landed in the staging buffer
    eph ^= 1;
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const uint32_t a = swz128(buf, t, j);
      const float4 x = lds_v4(a);
      sts_v4(a, x.x + sc * acc[4 * j], x.y + sc * acc[4 * j + 1], x.z + sc * acc[4 * j + 2], x.w + sc * acc[4 * j + 3]);
    }
    stage_store<false>(tmO, buf_ptr, col0, row0, lane);
  } else if constexpr (EPI == FC_EPI_MULAUX) {
    mbar_wait(ebar, eph);                       // saved gelu' tile has landed
    eph ^= 1;
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const uint32_t a = swz128(buf, t, j);
      const uint4 q = lds_u4(a);
      const float2 q0 = bf2_to_f2(q.x), q1 = bf2_to_f2(q.y), q2 = bf2_to_f2(q.z), q3 = bf2_to_f2(q.w);
      sts_u4(a, pack_bf16(acc[8 * j] * q0.x, acc[8 * j + 1] * q0.y), pack_bf16(acc[8 * j + 2] * q1.x, acc[8 * j + 3] * q1.y),
             pack_bf16(acc[8 * j + 4] * q2.x, acc[8 * j + 5] * q2.y), pack_bf16(acc[8 * j + 6] * q3.x, acc[8 * j + 7] * q3.y));
    }
    if (gp.colsum != nullptr) {                  // bias gradient: column sums of the bf16 values just staged
      __syncwarp();
      float s0 = 0.f, s1 = 0.f;                 // lane owns columns 2*lane, 2*lane+1 of the tile
#pragma unroll 8
      for (int r = 0; r < 32; ++r) {
        const float2 v = bf2_to_f2(lds_u32(swz128(buf, r, lane >> 2) + (lane & 3) * 4));
        s0 += v.x;
        s1 += v.y;
      }
      const int c = col0 + 2 * lane;
      if (c < p.N) {
        atomicAdd(gp.colsum + c, s0);
        atomicAdd(gp.colsum + c + 1, s1);
      }
    }
    stage_store<false>(tmO, buf_ptr, col0, row0, lane);
  } else if constexpr (EPI == FC_EPI_F32) {
    stage_acquire(lane);
#pragma unroll
    for (int j = 0; j < 8; ++j) sts_v4(swz128(buf, t, j), acc[4 * j], acc[4 * j + 1], acc[4 * j + 2], acc[4 * j + 3]);
    stage_store<false>(tmO, buf_ptr, col0, row0, lane);
  } else if constexpr (EPI == FC_EPI_ATOMIC_F32) {
    stage_acquire(lane);
#pragma unroll
    for (int j = 0; j < 8; ++j)
      sts_v4(swz128(buf, t, j), p.alpha * acc[4 * j], p.alpha * acc[4 * j + 1], p.alpha * acc[4 * j + 2], p.alpha * acc[4 * j + 3]);
    stage_store<true>(tmO, buf_ptr, col0, row0, lane);     // cp.reduce.async.bulk .add.f32: split-K / grad accumulate
  }
}

// Persistent, warp-specialised: each CTA (one per SM) walks tiles  t = blockIdx.x, +gridDim.x, ...
//   tile -> (split, m_blk, n_blk), n fastest so CTAs running side by side share the A rows in L2.
// smem ring (TMA -> MMA) runs across tile boundaries; the accumulator is double-buffered in TMEM so the
// epilogue of tile i overlaps the main loop of tile i+1.
template <int BN, int A_MN, int B_MN, int EPI, int PAIR>
__global__ void __launch_bounds__(gemm_threads_for(EPI), 1)
gemm_bf16_kernel(const __grid_constant__ AllMaps maps, const __grid_constant__ GemmParams p) {
  using C = Cfg<BN, EPI, PAIR>;
  constexpr int STAGES = C::STAGES;
  constexpr bool pair = PAIR != 0;
  constexpr int stage_bytes = C::STAGE_BYTES;
  constexpr int EPI_WARPS = C::EPI_WARPS;
  constexpr int EPI_STAGE_BYTES = C::EPI_STAGE_BYTES;
  extern __shared__ __align__(1024) uint8_t smem[];   // swizzled tiles need 1024-byte alignment (checked below)
  if (smem_u32(smem) & 1023) __trap();
  uint8_t* epi_stage = smem + STAGES * C::STAGE_BYTES;
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + STAGES * C::STAGE_BYTES + EPI_STAGE_BYTES);
  uint64_t* empty_bar = full_bar + STAGES;
  uint64_t* tmem_full_bar = empty_bar + STAGES;      // [2]
  uint64_t* tmem_empty_bar = tmem_full_bar + 2;      // [2]
  uint64_t* epi_bar = tmem_empty_bar + 2;            // [EPI_WARPS] one per epilogue warp (fused-operand TMA loads)
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(epi_bar + EPI_WARPS);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int total_kb = (p.K + BK - 1) / BK;
  const TileWalk<PAIR> walk(p);
  // pair mode: both CTAs use the same stage offsets; the leader's MMAs read A and B from both shared memories.

  if (warp == 0 && lane < p.groups) {
    const GroupMaps& gm = maps.g[lane];
    prefetch_tmap(&gm.a);
    prefetch_tmap(&gm.b);
    if (EPI != FC_EPI_PATCH) prefetch_tmap(&gm.o);
    if (EPI == FC_EPI_GELU) prefetch_tmap(&gm.o2);
    if (EpiTraits<EPI>::kLoads) prefetch_tmap(&gm.r);
    if (p.split) {
      prefetch_tmap(&gm.a_lo);
      prefetch_tmap(&gm.b_lo);
    }
  }
  if (warp == 1) {
    if (lane == 0) {
      for (int s = 0; s < STAGES; ++s) {
        mbar_init(&full_bar[s], 1);
        mbar_init(&empty_bar[s], 1);
      }
      for (int b = 0; b < 2; ++b) {
        mbar_init(&tmem_full_bar[b], 1);
        mbar_init(&tmem_empty_bar[b], pair ? 2 * EPI_WARPS : EPI_WARPS);    // one arrival per epilogue warp (of both CTAs)
      }
      for (int w = 0; w < EPI_WARPS; ++w) mbar_init(&epi_bar[w], 1);
      fence_barrier_init();
    }
    __syncwarp();
    if constexpr (pair) {
      tmem_alloc_pair(tmem_slot, 512);
    } else {
      tmem_alloc(tmem_slot, 512);
      tmem_relinquish();
    }
  }
  tc_fence_before();
  __syncthreads();
  if constexpr (pair) cluster_sync_all();           // the peer's barriers and TMEM exist before anything crosses CTAs
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  // Producer and MMA warps run their loops warp-uniform; only the TMA / MMA / commit / arrive instructions are
  // predicated on the elected lane.  (Inside an `if (lane == 0)` branch the compiler cannot keep descriptors in
  // uniform registers and wraps every UTCHMMA / UTMALDG in an ELECT + BRA.U.ANY loop: tools/mma_probe.cu.)
  if (warp == 0) {
    {
      const bool leader = elect_one();
      int s = 0;
      uint32_t ph = 0;
      for (int tile = walk.first; tile < walk.total; tile += walk.step) {
        int grp, split, m_blk, n_blk;
        walk.decode(tile, grp, split, m_blk, n_blk);
        const int m0 = m_blk * BM, n0 = n_blk * BN;
        const int kb0 = split * p.kb_per_split, kb1 = min(total_kb, kb0 + p.kb_per_split);
        // split-operand mode: three passes over the K range — (A_hi, B_hi), (A_hi, B_lo), (A_lo, B_hi) — into one accumulator
        for (int pass = 0; pass < (p.split ? 3 : 1); ++pass) {
        const CUtensorMap* tmA = pass == 2 ? &maps.g[grp].a_lo : &maps.g[grp].a;
        const CUtensorMap* tmB = pass == 1 ? &maps.g[grp].b_lo : &maps.g[grp].b;
        for (int kb = (p.debug & 2) ? kb1 : kb0; kb < kb1; ++kb) {
          mbar_wait(&empty_bar[s], ph ^ 1);
          uint8_t* a_dst = smem + s * stage_bytes;
          uint8_t* b_dst = a_dst + A_BYTES;
          if constexpr (pair) {
            // Both CTAs' loads complete their bytes on the LEADER's full barrier (its MMA warp is the only consumer);
            // the leader alone arms it, with the bytes of both — a local arrive: no cluster-scope release in this loop.
            // (The peer's bytes may land before the leader arms the phase: the transaction count just goes negative;
            //  they cannot run a phase ahead, because the peer's slot is only freed by the commit of this phase's MMAs.)
            const uint32_t lbar = mapa_u32(smem_u32(&full_bar[s]), 0);
            if (walk.rank == 0 && leader) mbar_arrive_expect_tx(&full_bar[s], 2 * stage_bytes);
            if (A_MN) {
              if (leader) tma_load_2d_pair(a_dst, tmA, lbar, m0, kb * BK);
              if (leader) tma_load_2d_pair(a_dst + 8192, tmA, lbar, m0 + 64, kb * BK);
            } else {
              if (leader) tma_load_2d_pair(a_dst, tmA, lbar, kb * BK, m0);
            }
            const int nh = n0 + walk.rank * (BN / 2);          // this CTA's half of the B tile
            if (B_MN) {
#pragma unroll
              for (int j = 0; j < BN / 128; ++j) if (leader) tma_load_2d_pair(b_dst + j * 8192, tmB, lbar, nh + j * 64, kb * BK);
            } else {      // box {64 (k), BN/2 (n)}
              if (leader) tma_load_2d_pair(b_dst, tmB, lbar, kb * BK, nh);
            }
          } else {
            if (leader) mbar_arrive_expect_tx(&full_bar[s], C::STAGE_BYTES);
            if (A_MN) {   // [K, M] global: boxes {64 (m), 64 (k)}
              if (leader) tma_load_2d(a_dst, tmA, &full_bar[s], m0, kb * BK);
              if (leader) tma_load_2d(a_dst + 8192, tmA, &full_bar[s], m0 + 64, kb * BK);
            } else {      // [M, K] global: box {64 (k), 128 (m)}
              if (leader) tma_load_2d(a_dst, tmA, &full_bar[s], kb * BK, m0);
            }
            if (B_MN) {
#pragma unroll
              for (int j = 0; j < BN / 64; ++j) if (leader) tma_load_2d(b_dst + j * 8192, tmB, &full_bar[s], n0 + j * 64, kb * BK);
            } else {      // box {64 (k), BN (n)}
              if (leader) tma_load_2d(b_dst, tmB, &full_bar[s], kb * BK, n0);
            }
          }
          if (++s == STAGES) { s = 0; ph ^= 1; }
        }
        }
      }
    }
  } else if (warp == 1) {
    if (!(pair && walk.rank != 0)) {     // pair mode: the leader CTA issues for both
      const bool leader = elect_one();
      constexpr uint32_t idesc = umma_idesc_bf16(pair ? 2 * BM : BM, BN, A_MN, B_MN);
      int s = 0;
      uint32_t ph = 0;
      int it = 0;
      for (int tile = walk.first; tile < walk.total; tile += walk.step, ++it) {
        const int split = (tile % walk.per_group) / walk.tiles_mn;
        const int kb0 = split * p.kb_per_split, kb1 = min(total_kb, kb0 + p.kb_per_split);
        const int buf = it & 1;
        mbar_wait(&tmem_empty_bar[buf], ((it >> 1) & 1) ^ 1);      // epilogue drained this accumulator
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + buf * TMEM_BUF_COLS;
        if ((p.debug & 2) && !pair) {          // measurement aid: no operands, no MMAs — epilogue-only timing
          if (leader) mbar_arrive(&tmem_full_bar[buf]);
          continue;
        }
        uint32_t acc_flag = 0;                 // 0 for the very first MMA of the tile, 1 afterwards
        for (int vkb = 0, nvkb = (kb1 - kb0) * (p.split ? 3 : 1); vkb < nvkb; ++vkb) {
          mbar_wait(&full_bar[s], ph);
          tc_fence_after();
          const uint32_t a_addr = smem_u32(smem + s * stage_bytes);
          const uint32_t b_addr = a_addr + A_BYTES;
#pragma unroll
          for (int k = 0; k < BK / 16; ++k) {
            // K-major: 16 K-elements = 32 B inside the 128-B swizzle row; SBO = 8 rows * 128 B.
            // MN-major: 16 K-rows of 128 B = 2048 B; LBO = next 64-wide MN block (8 KB), SBO = 8 K-rows.
            const uint64_t ad = A_MN ? umma_smem_desc(a_addr + k * 2048, 8192, 1024)
                                     : umma_smem_desc(a_addr + k * 32, 16, 1024);
            const uint64_t bd = B_MN ? umma_smem_desc(b_addr + k * 2048, 8192, 1024)
                                     : umma_smem_desc(b_addr + k * 32, 16, 1024);
            if (leader) {
              if constexpr (pair) umma_bf16_pair(d_tmem, ad, bd, idesc, acc_flag);
              else umma_bf16(d_tmem, ad, bd, idesc, acc_flag);
            }
            acc_flag = 1u;
          }
          if (leader) {
            if constexpr (pair) umma_commit_pair(&empty_bar[s]);   // slot free in both CTAs once these MMAs retire
            else umma_commit(&empty_bar[s]);     // smem slot free once these MMAs retire
          }
          if (++s == STAGES) { s = 0; ph ^= 1; }
        }
        if (leader) {
          if constexpr (pair) umma_commit_pair(&tmem_full_bar[buf]);   // accumulator halves complete in both CTAs
          else umma_commit(&tmem_full_bar[buf]); // accumulator complete
        }
      }
    }
  } else {
    constexpr int WPQ = EPI_WARPS / 4;         // warps per TMEM lane quarter
    const int q = warp & 3;                    // lane quarter this warp may access
    const int sub = (warp - 2) >> 2;           // which warp of that quarter: chunks c = sub, sub + WPQ, ...
    uint8_t* buf_ptr = epi_stage + (warp - 2) * EPI_BUF_BYTES;
    const uint32_t buf = smem_u32(buf_ptr);
    const uint32_t bias_smem = smem_u32(epi_stage) + EPI_WARPS * EPI_BUF_BYTES + (warp - 2) * EPI_BIAS_BYTES;
    uint64_t* ebar = &epi_bar[warp - 2];
    uint32_t eph = 0;
    constexpr int CW = EPI == FC_EPI_PATCH ? 32 : EpiTraits<EPI>::kCW;
    int it = 0;
    // pair mode: the accumulator is released to the LEADER's MMA warp by the epilogue warps of both CTAs
    uint32_t leader_empty[2] = {0u, 0u};
    if constexpr (pair) {
      leader_empty[0] = mapa_u32(smem_u32(&tmem_empty_bar[0]), 0);
      leader_empty[1] = mapa_u32(smem_u32(&tmem_empty_bar[1]), 0);
    }
    auto release_acc = [&](int b) {
      if constexpr (pair) mbar_arrive_cluster(leader_empty[b]);
      else mbar_arrive(&tmem_empty_bar[b]);
    };
    for (int tile = walk.first; tile < walk.total; tile += walk.step, ++it) {
      int grp, split_unused, m_blk, n_blk;
      walk.decode(tile, grp, split_unused, m_blk, n_blk);
      const GroupPtrs& gp = p.g[grp];
      const CUtensorMap* tmO = &maps.g[grp].o;
      const CUtensorMap* tmO2 = &maps.g[grp].o2;
      const CUtensorMap* tmR = &maps.g[grp].r;
      const int m0 = m_blk * BM, n0 = n_blk * BN;
      const int buf_i = it & 1;
      mbar_wait(&tmem_full_bar[buf_i], (it >> 1) & 1);
      tc_fence_after();
      const uint32_t taddr = tmem_base + buf_i * TMEM_BUF_COLS + (static_cast<uint32_t>(q * 32) << 16);
      const int row0 = m0 + q * 32;
      // chunks this warp owns: c = sub, sub+WPQ, ... ; the last one it will actually read (col0 < N, rows < M)
      int last_c = -1;
      if (row0 < p.M)
        for (int c = sub; c < BN / CW; c += WPQ)
          if (n0 + c * CW < p.N) last_c = c;
      if (last_c < 0 || (p.debug & 1)) {       // nothing to read in this tile (or main-loop-only timing)
        __syncwarp();
        if (lane == 0) release_acc(buf_i);
        continue;
      }
#pragma unroll 1
      for (int c = sub; c <= last_c; c += WPQ) {
        const int col0 = n0 + c * CW;
        if constexpr (EPI == FC_EPI_PATCH) {
          // legacy path (row remap b*P+t -> b*(P+1)+1+t cannot be expressed as one TMA box): transpose via smem
          float4 bias = make_float4(0.f, 0.f, 0.f, 0.f);
          const int bcol = col0 + (lane & 7) * 4;
          if (gp.bias != nullptr && bcol < p.N) bias = __ldg(reinterpret_cast<const float4*>(gp.bias + bcol));
          EpiPre<EPI> pre;
          float acc[32];
          tmem_ld_32x32(taddr + c * 32, acc);
          tmem_ld_wait();
          if (c == last_c) {
            tc_fence_before();
            __syncwarp();
            if (lane == 0) release_acc(buf_i);
          }
          if (lane < 16) {
#pragma unroll
            for (int j = 0; j < 8; ++j)
              sts_v4(buf + (lane * STAGE_LD + j * 4) * 4, acc[4 * j], acc[4 * j + 1], acc[4 * j + 2], acc[4 * j + 3]);
          }
          __syncwarp();
          epilogue_block<EPI, 0>(p, gp, buf, row0, col0, lane, bias, pre);
          __syncwarp();
          if (lane >= 16) {
#pragma unroll
            for (int j = 0; j < 8; ++j)
              sts_v4(buf + ((lane - 16) * STAGE_LD + j * 4) * 4, acc[4 * j], acc[4 * j + 1], acc[4 * j + 2], acc[4 * j + 3]);
          }
          __syncwarp();
          epilogue_block<EPI, 1>(p, gp, buf, row0, col0, lane, bias, pre);
          __syncwarp();
        } else {
          if constexpr (EpiTraits<EPI>::kLoads) {     // fused operand tile (residual / saved gelu') -> staging buffer
            if (elect_one()) {                 // (the same lane as in stage_store / stage_acquire)
              tma_store_wait_read();           // the previous store out of this buffer has drained
              mbar_arrive_expect_tx(ebar, EPI_BUF_BYTES);
              tma_load_2d(buf_ptr, tmR, ebar, col0, row0);
            }
          }
          // the chunk's bias: lane l fetches columns col0 + 2l, 2l+1 now (latency overlaps the TMEM read) ...
          float2 b2 = make_float2(0.f, 0.f);
          if (gp.bias != nullptr && 2 * lane < CW && col0 + 2 * lane < p.N)
            b2 = __ldg(reinterpret_cast<const float2*>(gp.bias + col0 + 2 * lane));
          float acc[CW];
          {
            float(&lo)[32] = *reinterpret_cast<float(*)[32]>(&acc[0]);
            tmem_ld_32x32(taddr + c * CW, lo);
            if constexpr (CW == 64) {
              float(&hi)[32] = *reinterpret_cast<float(*)[32]>(&acc[32]);
              tmem_ld_32x32(taddr + c * CW + 32, hi);
            }
          }
          tmem_ld_wait();
          if (c == last_c) {                   // last TMEM read of this tile by this warp: hand the buffer back
            tc_fence_before();
            __syncwarp();
            if (lane == 0) release_acc(buf_i);
          }
          if (gp.bias != nullptr) {             // ... and shares it with the other lanes through shared memory
            if (2 * lane < CW)
              asm volatile("st.shared.v2.f32 [%0], {%1, %2};" ::"r"(bias_smem + 8 * lane), "f"(b2.x), "f"(b2.y) : "memory");
            __syncwarp();
          }
          epilogue_tma_chunk<EPI>(p, gp, tmO, tmO2, buf_ptr, buf, bias_smem, ebar, eph, acc, row0, col0, lane);
        }
      }
    }
    if (lane == 0) tma_store_wait_all();       // all bulk stores of this warp are globally complete before exit
  }
  __syncthreads();
  if constexpr (pair) cluster_sync_all();            // the peer may still read this CTA's smem / arrive on its barriers
  if (warp == 1) {
    tc_fence_after();
    if constexpr (pair) tmem_dealloc_pair(tmem_base, 512);
    else tmem_dealloc(tmem_base, 512);
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

// row-major matrix [rows, cols] (bf16, or fp32 when f32 != 0) with leading dimension ld; box = {box_cols, box_rows}
int make_tmap(CUtensorMap* m, const void* ptr, long long rows, long long cols, long long ld, int box_cols,
              int box_rows, int f32 = 0) {
  EncodeTiledFn fn = get_encode_fn();
  if (!fn) FC_FAIL(FC_ERR_CUDA, "cuTensorMapEncodeTiled entry point unavailable");
  cuuint64_t gdim[2] = {static_cast<cuuint64_t>(cols), static_cast<cuuint64_t>(rows)};
  cuuint64_t gstr[1] = {static_cast<cuuint64_t>(ld) * (f32 ? 4 : 2)};
  cuuint32_t box[2] = {static_cast<cuuint32_t>(box_cols), static_cast<cuuint32_t>(box_rows)};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = fn(m, f32 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT32 : CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(ptr), gdim, gstr, box, estr,
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

template <int BN, int A_MN, int B_MN, int EPI, int PAIR>
int launch(const AllMaps& maps, const GemmParams& p, int device, cudaStream_t st) {
  using C = Cfg<BN, EPI, PAIR>;
  auto kern = gemm_bf16_kernel<BN, A_MN, B_MN, EPI, PAIR>;
  FC_SMEM_OPT_IN(kern, (C::SMEM_BYTES));
  int grid = fc_num_sms(device);
  if (PAIR) {
    const int pair_tiles = ((p.m_tiles + 1) / 2) * p.n_tiles * p.splits * p.groups;
    grid &= ~1;
    if (grid > 2 * pair_tiles) grid = 2 * pair_tiles;
    grid = fc_apply_grid_cap(grid) & ~1;
    if (grid < 2) grid = 2;
  } else {
    const int total_tiles = p.m_tiles * p.n_tiles * p.splits * p.groups;
    if (grid > total_tiles) grid = total_tiles;
    grid = fc_apply_grid_cap(grid);
  }
  ProfRec rec{nullptr, nullptr, 2.0 * p.M * (double)p.N * p.K * p.groups};
  const bool prof = __atomic_load_n(&g_prof_on, __ATOMIC_RELAXED) != 0;
  if (prof) {
    cudaEventCreate(&rec.a);
    cudaEventCreate(&rec.b);
    cudaEventRecord(rec.a, st);
  }
  if (PAIR) {       // CTA pairs: clusters of two along x (cta_group::2 pairs rank r with r ^ 1)
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(grid);
    cfg.blockDim = dim3(gemm_threads_for(EPI));
    cfg.dynamicSmemBytes = C::SMEM_BYTES;
    cfg.stream = st;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = 2;
    attr[0].val.clusterDim.y = 1;
    attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    FC_CUDA_CHECK(cudaLaunchKernelEx(&cfg, kern, maps, p));
  } else {
    kern<<<grid, gemm_threads_for(EPI), C::SMEM_BYTES, st>>>(maps, p);
  }
  if (prof) {
    cudaEventRecord(rec.b, st);
    std::lock_guard<std::mutex> lk(g_prof_mu);
    g_prof.push_back(rec);
  }
  FC_LAUNCH_CHECK();
  return FC_OK;
}

// Only the (operand majors x epilogue) combinations the round uses are instantiated:
//   forward  A,B K-major        : BF16, GELU, RESID, F32, PATCH
//   dX       A K-major, B MN    : BF16, MULAUX, F32
//   dW       A,B MN-major       : ATOMIC_F32, F32
//   (A MN, B K)                 : F32 (parity tests)
// each as a single-CTA kernel (128 x BN tiles) and — except PATCH, and BN = 192 with an MN-major B whose half tile is
// not a whole number of 64-column blocks — as a CTA-pair kernel (cta_group::2, 256 x BN tiles).
template <int BN, int PAIR>
int launch_bn(const AllMaps& maps, const GemmParams& p, int a_mn, int b_mn, int device, cudaStream_t st) {
#define FC_CASE(AM, BMJ, E) \
  if (a_mn == AM && b_mn == BMJ && p.epi == E) return launch<BN, AM, BMJ, E, PAIR>(maps, p, device, st)
  FC_CASE(0, 0, FC_EPI_BF16); FC_CASE(0, 0, FC_EPI_GELU); FC_CASE(0, 0, FC_EPI_RESID); FC_CASE(0, 0, FC_EPI_F32);
  if constexpr (!PAIR) { FC_CASE(0, 0, FC_EPI_PATCH); }
  if constexpr (!PAIR || BN % 128 == 0) {
    FC_CASE(0, 1, FC_EPI_BF16); FC_CASE(0, 1, FC_EPI_MULAUX); FC_CASE(0, 1, FC_EPI_F32);
    FC_CASE(1, 1, FC_EPI_ATOMIC_F32); FC_CASE(1, 1, FC_EPI_F32);
  }
  FC_CASE(1, 0, FC_EPI_F32);
#undef FC_CASE
  FC_FAIL(FC_ERR_UNSUPPORTED, "fc_gemm_bf16: epilogue %d is not built for operand majors (a_mn=%d, b_mn=%d)", p.epi,
          a_mn, b_mn);
}

// Tile width: minimise (waves over the SMs) x (per-tile cost ~ BN + fixed overhead), i.e. trade the better
// operand reuse of wide tiles against wave quantisation and zero-padded columns.
int pick_bn(int M, int N, int splits, int sms) {      // splits: K splits x client groups
  const int cands[3] = {256, 192, 128};
  int best = 128;
  double best_cost = 1e30;
  const int m_tiles = (M + BM - 1) / BM;
  for (int bn : cands) {
    const long long tiles = (long long)m_tiles * ((N + bn - 1) / bn) * splits;
    const long long waves = (tiles + sms - 1) / sms;
    const double cost = (double)waves * (bn + 96.0);
    if (cost < best_cost - 1e-9) { best_cost = cost; best = bn; }
  }
  return best;
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

namespace {
int gemm_grouped_impl(int groups, int M, int N, int K, const void* const* A, const void* const* A_lo, long long lda,
                      int a_mn_major, const void* const* B, const void* const* B_lo, long long ldb, int b_mn_major, int epi,
                      void* const* out, void* const* out2, long long ldo, const float* const* bias,
                      const float* const* resid, const float* const* row_scale, int rows_per_group,
                      const void* const* aux, const float* const* pos, int patches, float alpha, int splits,
                      float* const* colsum, int device, void* stream) {
  FC_REQUIRE(groups >= 1 && groups <= FC_GEMM_MAX_GROUPS, "fc_gemm_bf16: %d groups (1..%d)", groups, FC_GEMM_MAX_GROUPS);
  FC_REQUIRE(M > 0 && N > 0 && K > 0, "fc_gemm_bf16: empty problem %d %d %d", M, N, K);
  FC_REQUIRE(N % 8 == 0 && lda % 8 == 0 && ldb % 8 == 0 && ldo % 4 == 0, "fc_gemm_bf16: N, lda, ldb must be multiples of 8");
  FC_REQUIRE(epi >= FC_EPI_BF16 && epi <= FC_EPI_PATCH, "fc_gemm_bf16: bad epilogue %d", epi);
  FC_REQUIRE(A != nullptr && B != nullptr && out != nullptr, "fc_gemm_bf16: null operand table");
  FC_REQUIRE((A_lo == nullptr) == (B_lo == nullptr), "fc_gemm_bf16: split operands need both low-order tables");
  auto at = [](auto tbl, int g) { return tbl ? tbl[g] : nullptr; };
  for (int g = 0; g < groups; ++g) {
    FC_REQUIRE(A[g] != nullptr && B[g] != nullptr && out[g] != nullptr, "fc_gemm_bf16: null operand / output (group %d)", g);
    FC_REQUIRE(epi != FC_EPI_GELU || at(out2, g) != nullptr, "fc_gemm_bf16: GELU epilogue needs out2");
    FC_REQUIRE(epi != FC_EPI_RESID || at(resid, g) != nullptr, "fc_gemm_bf16: RESID epilogue needs resid");
    FC_REQUIRE(epi != FC_EPI_MULAUX || at(aux, g) != nullptr, "fc_gemm_bf16: MULAUX epilogue needs aux");
    FC_REQUIRE(at(colsum, g) == nullptr || epi == FC_EPI_MULAUX, "fc_gemm_bf16: colsum is only fused into the MULAUX epilogue");
    FC_REQUIRE(epi != FC_EPI_PATCH || (at(pos, g) != nullptr && patches > 0), "fc_gemm_bf16: PATCH epilogue needs pos");
    // the epilogue branches on these once per launch: all groups must agree on which optional operands exist
    FC_REQUIRE((at(bias, g) == nullptr) == (at(bias, 0) == nullptr) && (at(row_scale, g) == nullptr) == (at(row_scale, 0) == nullptr) &&
               (at(colsum, g) == nullptr) == (at(colsum, 0) == nullptr), "fc_gemm_bf16: groups disagree on optional operands");
  }
  FC_REQUIRE(at(row_scale, 0) == nullptr || rows_per_group > 0, "fc_gemm_bf16: rows_per_group");
  FcDeviceGuard guard(device);
  const int total_kb = (K + BK - 1) / BK;
  const int sms = fc_num_sms(device);
  int bn = 0;
  if (splits <= 0 && epi == FC_EPI_ATOMIC_F32) {
    // auto split-K: (tile width, #splits) minimising  waves x (main loop + atomic epilogue) over the SMs
    double best = 1e30;
    const int cands[3] = {256, 192, 128};
    const int max_s = total_kb / 2 > 1 ? (total_kb / 2 < 64 ? total_kb / 2 : 64) : 1;
    for (int c : cands) {
      const long long base = (long long)((M + BM - 1) / BM) * ((N + c - 1) / c) * groups;
      for (int sp = 1; sp <= max_s; ++sp) {
        const int per = (total_kb + sp - 1) / sp;
        const long long tiles = base * ((total_kb + per - 1) / per);
        const long long waves = (tiles + sms - 1) / sms;
        const double cost = (double)waves * (per * (c + 64.0) + 2.0 * c + 200.0);
        if (cost < best - 1e-9) { best = cost; bn = c; splits = sp; }
      }
    }
  }
  if (splits < 1) splits = 1;
  if (splits > total_kb) splits = total_kb;
  FC_REQUIRE(splits == 1 || epi == FC_EPI_ATOMIC_F32, "fc_gemm_bf16: split-K needs the atomic epilogue");
  GemmParams p;
  memset(&p, 0, sizeof(p));
  p.M = M; p.N = N; p.K = K;
  p.groups = groups;
  p.kb_per_split = (total_kb + splits - 1) / splits;
  p.splits = (total_kb + p.kb_per_split - 1) / p.kb_per_split;
  if (bn == 0) bn = pick_bn(M, N, p.splits * groups, sms);
  p.m_tiles = (M + BM - 1) / BM;
  p.n_tiles = (N + bn - 1) / bn;
  p.epi = epi; p.ldo = static_cast<int>(ldo);
  p.rows_per_group = rows_per_group > 0 ? rows_per_group : 1;
  p.patches = patches; p.alpha = alpha;
  for (int g = 0; g < groups; ++g) {
    GroupPtrs& q = p.g[g];
    q.out = out[g]; q.out2 = at(out2, g); q.bias = at(bias, g); q.resid = at(resid, g); q.row_scale = at(row_scale, g);
    q.aux = reinterpret_cast<const __nv_bfloat16*>(at(aux, g)); q.pos = at(pos, g); q.colsum = at(colsum, g);
  }
  {
    static const int dbg = getenv("FC_GEMM_DEBUG") ? atoi(getenv("FC_GEMM_DEBUG")) : 0;
    p.debug = dbg;
    // Experiment knobs: FC_GEMM_BN forces the tile width; FC_GEMM_PAIR = 0 never / 1 whenever legal: CTA pairs
    // (cta_group::2, 256 x BN pair tiles).  Legal: not the PATCH epilogue; an MN-major B needs BN % 128 == 0 (each
    // CTA's half tile must be whole 64-column blocks).
    static const int force_bn = getenv("FC_GEMM_BN") ? atoi(getenv("FC_GEMM_BN")) : 0;
    static const int use_pairs = getenv("FC_GEMM_PAIR") ? atoi(getenv("FC_GEMM_PAIR")) : 0;
    if (force_bn == 128 || force_bn == 192 || force_bn == 256) {
      bn = force_bn;
      p.n_tiles = (N + bn - 1) / bn;
    }
    p.split = A_lo != nullptr ? 1 : 0;
    p.pair = (use_pairs && !p.split && epi != FC_EPI_PATCH && (!b_mn_major || bn % 128 == 0) && !(p.debug & 2)) ? 1 : 0;
  }
  AllMaps maps;
  memset(&maps, 0, sizeof(maps));
  const int out16 = (epi == FC_EPI_BF16 || epi == FC_EPI_GELU || epi == FC_EPI_MULAUX);
  for (int g = 0; g < groups; ++g) {
    GroupMaps& gm = maps.g[g];
    int rc;
    // K-major operand: global [rows, K]; MN-major operand: global [K, rows].
    rc = a_mn_major ? make_tmap(&gm.a, A[g], K, M, lda, 64, 64) : make_tmap(&gm.a, A[g], M, K, lda, 64, BM);
    if (rc) return rc;
    rc = b_mn_major ? make_tmap(&gm.b, B[g], K, N, ldb, 64, 64) : make_tmap(&gm.b, B[g], N, K, ldb, 64, p.pair ? bn / 2 : bn);
    if (rc) return rc;
    if (p.split) {
      FC_REQUIRE(A_lo[g] != nullptr && B_lo[g] != nullptr, "fc_gemm_bf16: null low-order operand (group %d)", g);
      rc = a_mn_major ? make_tmap(&gm.a_lo, A_lo[g], K, M, lda, 64, 64) : make_tmap(&gm.a_lo, A_lo[g], M, K, lda, 64, BM);
      if (rc) return rc;
      rc = b_mn_major ? make_tmap(&gm.b_lo, B_lo[g], K, N, ldb, 64, 64) : make_tmap(&gm.b_lo, B_lo[g], N, K, ldb, 64, bn);
      if (rc) return rc;
    }
    // epilogue tiles: 32 rows x 128 bytes (64 bf16 / 32 fp32 columns), stored / reduced / loaded by TMA
    if (epi != FC_EPI_PATCH) {
      FC_REQUIRE((reinterpret_cast<uintptr_t>(out[g]) & 15) == 0 && (ldo * (out16 ? 2 : 4)) % 16 == 0,
                 "fc_gemm_bf16: output must be 16-byte aligned with a 16-byte multiple row pitch");
      rc = make_tmap(&gm.o, out[g], M, N, ldo, out16 ? 64 : 32, 32, !out16);
      if (rc) return rc;
      if (epi == FC_EPI_GELU) {
        rc = make_tmap(&gm.o2, out2[g], M, N, ldo, 64, 32, 0);
        if (rc) return rc;
      }
      if (epi == FC_EPI_RESID) {
        rc = make_tmap(&gm.r, resid[g], M, N, ldo, 32, 32, 1);
        if (rc) return rc;
      }
      if (epi == FC_EPI_MULAUX) {
        rc = make_tmap(&gm.r, aux[g], M, N, ldo, 64, 32, 0);
        if (rc) return rc;
      }
    }
  }
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  const int am = a_mn_major ? 1 : 0, bm = b_mn_major ? 1 : 0;
  if (p.pair) {
    if (bn == 256) return launch_bn<256, 1>(maps, p, am, bm, device, st);
    if (bn == 192) return launch_bn<192, 1>(maps, p, am, bm, device, st);
    return launch_bn<128, 1>(maps, p, am, bm, device, st);
  }
  if (bn == 256) return launch_bn<256, 0>(maps, p, am, bm, device, st);
  if (bn == 192) return launch_bn<192, 0>(maps, p, am, bm, device, st);
  return launch_bn<128, 0>(maps, p, am, bm, device, st);
}
}  // namespace

extern "C" int fc_gemm_bf16_grouped(int groups, int M, int N, int K, const void* const* A, long long lda, int a_mn_major,
                                    const void* const* B, long long ldb, int b_mn_major, int epi, void* const* out,
                                    void* const* out2, long long ldo, const float* const* bias,
                                    const float* const* resid, const float* const* row_scale, int rows_per_group,
                                    const void* const* aux, const float* const* pos, int patches, float alpha,
                                    int splits, float* const* colsum, int device, void* stream) {
  return gemm_grouped_impl(groups, M, N, K, A, nullptr, lda, a_mn_major, B, nullptr, ldb, b_mn_major, epi, out, out2, ldo, bias,
                           resid, row_scale, rows_per_group, aux, pos, patches, alpha, splits, colsum, device, stream);
}

// fp32-accurate GEMM on the bf16 tensor pipe: each operand is a (hi, lo) pair of bf16 matrices, X = X_hi + X_lo with
// X_lo = bf16(X - X_hi) (16 mantissa bits), and the accumulator receives A_hi B_hi + A_hi B_lo + A_lo B_hi — relative
// error ~2^-16 per product, against 2^-8 for plain bf16 operands and 2^-10 for tcgen05's kind::tf32.  Used by the
// validation mode (precision = 'fp32'); only the fp32-output epilogues make sense with it.
extern "C" int fc_gemm_split(int M, int N, int K, const void* A_hi, const void* A_lo, long long lda, int a_mn_major,
                             const void* B_hi, const void* B_lo, long long ldb, int b_mn_major, int epi, void* out,
                             long long ldo, const float* bias, const float* resid, const float* row_scale,
                             int rows_per_group, const float* pos, int patches, float alpha, int splits, int device,
                             void* stream) {
  FC_REQUIRE(epi == FC_EPI_F32 || epi == FC_EPI_RESID || epi == FC_EPI_ATOMIC_F32 || epi == FC_EPI_PATCH,
             "fc_gemm_split: only the fp32-output epilogues (F32, RESID, ATOMIC_F32, PATCH)");
  FC_REQUIRE(A_hi && A_lo && B_hi && B_lo && out, "fc_gemm_split: null operand");
  void* out2 = nullptr;
  const void* aux = nullptr;
  float* colsum = nullptr;
  return gemm_grouped_impl(1, M, N, K, &A_hi, &A_lo, lda, a_mn_major, &B_hi, &B_lo, ldb, b_mn_major, epi, &out, &out2, ldo,
                           &bias, &resid, &row_scale, rows_per_group, &aux, &pos, patches, alpha, splits, &colsum, device,
                           stream);
}

// One operand set: the plain GEMM (ref: every F.linear of the reference's Block).
extern "C" int fc_gemm_bf16(int M, int N, int K, const void* A, long long lda, int a_mn_major, const void* B,
                            long long ldb, int b_mn_major, int epi, void* out, void* out2, long long ldo,
                            const float* bias, const float* resid, const float* row_scale, int rows_per_group,
                            const void* aux, const float* pos, int patches, float alpha, int splits,
                            float* colsum, int device, void* stream) {
  FC_REQUIRE(M > 0 && N > 0 && K > 0, "fc_gemm_bf16: empty problem %d %d %d", M, N, K);
  FC_REQUIRE(out != nullptr, "fc_gemm_bf16: null output");
  return fc_gemm_bf16_grouped(1, M, N, K, &A, lda, a_mn_major, &B, ldb, b_mn_major, epi, &out, &out2, ldo, &bias, &resid,
                              &row_scale, rows_per_group, &aux, &pos, patches, alpha, splits, &colsum, device, stream);
}
