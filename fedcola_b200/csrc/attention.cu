// Fused short-sequence attention, forward and backward, head_dim = 64 (every reference factory:
// 192/3, 384/6, 768/12 — /root/reference/src/models/mome.py:937-1030).
//
// Replaces Attention.forward between the qkv and proj Linears (mome.py:153-165): q*scale, fp32 QK^T,
// fp32 softmax, cast, P·V, head merge — and its autograd.  No attention mask exists in the reference
// (pad tokens are attended); only the padding rows/columns this kernel adds to reach a
// multiple of 16 are masked.  One CTA per (sample, head); the whole head (N <= 256 tokens) lives in
// shared memory, so the [B,H,N,N] probability tensor is never materialised in HBM.
//
// This file holds the BACKWARD: warp-level mma.sync.m16n8k16 (bf16 in, fp32 accumulate) with ldmatrix from
// swizzled shared memory (round-1 implementation; its tcgen05 rewrite is the next step).  The forward is the
// tcgen05/TMEM kernel in attention_tc.cu.
#include "common.cuh"
#include "../../include/fedcola_b200.h"

namespace {

constexpr int HD = 64;        // head dim
constexpr int LDS = 64;       // smem row = 128 B = 8 chunks of 16 B, XOR-swizzled by (row & 7): conflict-free ldmatrix
                              // without padding, so the backward kernel's 4 staged matrices fit twice per SM

// element offset of 16-byte chunk `chunk` (0..7) of row r
__device__ __forceinline__ int swz(int r, int chunk) { return r * LDS + ((chunk ^ (r & 7)) << 3); }

__device__ __forceinline__ uint32_t smem_addr(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void ldsm_x4(uint32_t (&r)[4], const __nv_bfloat16* p) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.shared.b16 {%0,%1,%2,%3}, [%4];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]) : "r"(smem_addr(p)));
}
__device__ __forceinline__ void ldsm_x4_t(uint32_t (&r)[4], const __nv_bfloat16* p) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.trans.shared.b16 {%0,%1,%2,%3}, [%4];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]) : "r"(smem_addr(p)));
}
// D(16x8, f32) += A(16x16, bf16, row) * B(16x8, bf16, col)
__device__ __forceinline__ void mma16816(float (&d)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
  asm volatile(
      "mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
      : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3])
      : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}
__device__ __forceinline__ uint32_t pack2(float a, float b) {
  __nv_bfloat162 v = __floats2bfloat162_rn(a, b);
  return *reinterpret_cast<uint32_t*>(&v);
}

// A-operand fragments of a 16-row tile (rows r0..r0+15) for all 4 k16 steps of the 64-wide head dim.
__device__ __forceinline__ void load_a_frags(uint32_t (&a)[4][4], const __nv_bfloat16* s, int r0, int lane) {
#pragma unroll
  for (int kk = 0; kk < 4; ++kk) ldsm_x4(a[kk], s + swz(r0 + (lane & 15), kk * 2 + (lane >> 4)));
}
// acc(16 x 16 cols [c0, c0+16)) = A(16 x 64) * M[c0..c0+16, 0..64)^T — M rows are the "n" index, head dim is k.
__device__ __forceinline__ void mma_rowsT(float (&acc)[2][4], const uint32_t (&a)[4][4], const __nv_bfloat16* m, int c0,
                                          int lane) {
#pragma unroll
  for (int j = 0; j < 2; ++j)
#pragma unroll
    for (int i = 0; i < 4; ++i) acc[j][i] = 0.f;
#pragma unroll
  for (int j = 0; j < 2; ++j) {
#pragma unroll
    for (int kk = 0; kk < 4; kk += 2) {
      uint32_t b[4];   // (n-tile j) x (k steps kk, kk+1)
      ldsm_x4(b, m + swz(c0 + j * 8 + (lane & 7), kk * 2 + (lane >> 3)));
      mma16816(acc[j], a[kk], b[0], b[1]);
      mma16816(acc[j], a[kk + 1], b[2], b[3]);
    }
  }
}
// out(16 x 64) += P(16 x 16, as A fragments) * M[r0..r0+16, 0..64)  — M rows are the k index (needs .trans)
__device__ __forceinline__ void mma_rows(float (&out)[8][4], const uint32_t (&p)[4], const __nv_bfloat16* m, int r0,
                                         int lane) {
#pragma unroll
  for (int n = 0; n < 8; n += 2) {
    uint32_t b[4];
    ldsm_x4_t(b, m + swz(r0 + (lane & 7) + ((lane >> 3) & 1) * 8, n + (lane >> 4)));
    mma16816(out[n], p, b[0], b[1]);
    mma16816(out[n + 1], p, b[2], b[3]);
  }
}

__device__ __forceinline__ void stage_rows(__nv_bfloat16* dst, const __nv_bfloat16* src, long long row_stride, int n,
                                           int npad) {
  // 64 bf16 per row = 8 x 16 B; zero-fill padding rows
  for (int i = threadIdx.x; i < npad * 8; i += blockDim.x) {
    const int r = i >> 3, c = i & 7;
    uint4 v = make_uint4(0, 0, 0, 0);
    if (r < n) v = *reinterpret_cast<const uint4*>(src + (size_t)r * row_stride + c * 8);
    *reinterpret_cast<uint4*>(dst + swz(r, c)) = v;
  }
}

// Column sums over the 16 rows of a warp tile of two packed bf16 pairs (rows g and g+8, columns 2t, 2t+1):
// reduce over the 8 row groups with shuffles, then lanes 0..3 add into shared memory.
__device__ __forceinline__ void tile_colsum(float* dst, uint32_t w0, uint32_t w1, int lane) {
  const float2 a = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&w0));
  const float2 b = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&w1));
  float x = a.x + b.x, y = a.y + b.y;
#pragma unroll
  for (int o = 4; o <= 16; o <<= 1) {
    x += __shfl_xor_sync(0xffffffffu, x, o);
    y += __shfl_xor_sync(0xffffffffu, y, o);
  }
  if (lane < 4) {
    atomicAdd(dst, x);
    atomicAdd(dst + 1, y);
  }
}

// ---- backward -----------------------------------------------------------------------------------
// dV = P^T dO;  dP = dO V^T;  dS = P ⊙ (dP - D), D_i = sum_c dO_ic O_ic;  dQ = scale dS K;  dK = scale dS^T Q
template <int NPAD, int NW>
__global__ void __launch_bounds__(NW * 32, 2) attn_bwd_kernel(const __nv_bfloat16* __restrict__ qkv,
                                                       const __nv_bfloat16* __restrict__ o_fwd,
                                                       const __nv_bfloat16* __restrict__ d_out,
                                                       const float* __restrict__ lse_g,
                                                       __nv_bfloat16* __restrict__ dqkv, float* __restrict__ dbias,
                                                       int N, int H, float scale) {
  extern __shared__ __align__(16) uint8_t smem_raw[];
  __nv_bfloat16* Qs = reinterpret_cast<__nv_bfloat16*>(smem_raw);
  __nv_bfloat16* Ks = Qs + NPAD * LDS;
  __nv_bfloat16* Vs = Ks + NPAD * LDS;
  __nv_bfloat16* Gs = Vs + NPAD * LDS;                       // dO
  float* lse = reinterpret_cast<float*>(Gs + NPAD * LDS);    // [NPAD]
  float* Dr = lse + NPAD;                                    // [NPAD]
  float* s_db = Dr + NPAD;                                   // [3*64] column sums of dQ | dK | dV of this head
  const int b = blockIdx.x / H, h = blockIdx.x % H;
  const int d = H * HD;
  const long long rs = 3LL * d;
  const __nv_bfloat16* base = qkv + (size_t)b * N * rs + h * HD;
  const __nv_bfloat16* gb = d_out + (size_t)b * N * d + h * HD;
  const __nv_bfloat16* ofb = o_fwd + (size_t)b * N * d + h * HD;
  stage_rows(Qs, base, rs, N, NPAD);
  stage_rows(Ks, base + d, rs, N, NPAD);
  stage_rows(Vs, base + 2 * d, rs, N, NPAD);
  stage_rows(Gs, gb, d, N, NPAD);
  for (int r = threadIdx.x; r < NPAD; r += blockDim.x) lse[r] = r < N ? lse_g[((size_t)b * H + h) * N + r] : INFINITY;
  for (int i = threadIdx.x; i < 3 * HD; i += blockDim.x) s_db[i] = 0.f;
  __syncthreads();
  // D_i = <dO_i, O_i>: 8 lanes per row, 8 elements each
  for (int i = threadIdx.x; i < NPAD * 8; i += blockDim.x) {
    const int r = i >> 3, c = i & 7;
    float acc = 0.f;
    if (r < N) {
      const uint4 ov = *reinterpret_cast<const uint4*>(ofb + (size_t)r * d + c * 8);
      const uint4 gv = *reinterpret_cast<const uint4*>(Gs + swz(r, c));
      const __nv_bfloat162* o2 = reinterpret_cast<const __nv_bfloat162*>(&ov);
      const __nv_bfloat162* g2 = reinterpret_cast<const __nv_bfloat162*>(&gv);
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        const float2 a = __bfloat1622float2(o2[k]), e = __bfloat1622float2(g2[k]);
        acc += a.x * e.x + a.y * e.y;
      }
    }
    acc += __shfl_xor_sync(0xffffffffu, acc, 1);
    acc += __shfl_xor_sync(0xffffffffu, acc, 2);
    acc += __shfl_xor_sync(0xffffffffu, acc, 4);
    if (c == 0) Dr[r] = acc;
  }
  __syncthreads();

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, g = lane >> 2, t = lane & 3;
  constexpr int NT = NPAD / 16;
  const int nwarps = blockDim.x >> 5;
  __nv_bfloat16* dq_base = dqkv + (size_t)b * N * rs + h * HD;

  // ---- pass A: query tiles -> dQ ----
  for (int mt = warp; mt < NT; mt += nwarps) {
    uint32_t qa[4][4], ga[4][4];
    load_a_frags(qa, Qs, mt * 16, lane);
    load_a_frags(ga, Gs, mt * 16, lane);
    const int row0 = mt * 16 + g, row1 = row0 + 8;
    const float l0 = lse[row0], l1 = lse[row1], d0 = Dr[row0], d1 = Dr[row1];
    float dq[8][4];
#pragma unroll
    for (int n = 0; n < 8; ++n)
#pragma unroll
      for (int i = 0; i < 4; ++i) dq[n][i] = 0.f;
#pragma unroll 1
    for (int ct = 0; ct < NT; ++ct) {
      float s[2][4], dp[2][4];
      mma_rowsT(s, qa, Ks, ct * 16, lane);
      mma_rowsT(dp, ga, Vs, ct * 16, lane);
      uint32_t ds[4];
      float v[2][4];
#pragma unroll
      for (int j = 0; j < 2; ++j)
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          const int col = ct * 16 + j * 8 + 2 * t + (i & 1);
          const float p = col < N ? __expf(s[j][i] * scale - (i < 2 ? l0 : l1)) : 0.f;
          v[j][i] = p * (dp[j][i] - (i < 2 ? d0 : d1));
        }
      ds[0] = pack2(v[0][0], v[0][1]); ds[1] = pack2(v[0][2], v[0][3]);
      ds[2] = pack2(v[1][0], v[1][1]); ds[3] = pack2(v[1][2], v[1][3]);
      mma_rows(dq, ds, Ks, ct * 16, lane);
    }
#pragma unroll
    for (int n = 0; n < 8; ++n) {
      const uint32_t w0 = pack2(dq[n][0] * scale, dq[n][1] * scale), w1 = pack2(dq[n][2] * scale, dq[n][3] * scale);
      if (row0 < N) *reinterpret_cast<uint32_t*>(dq_base + (size_t)row0 * rs + n * 8 + 2 * t) = w0;
      if (row1 < N) *reinterpret_cast<uint32_t*>(dq_base + (size_t)row1 * rs + n * 8 + 2 * t) = w1;
      if (dbias != nullptr) tile_colsum(s_db + n * 8 + 2 * t, row0 < N ? w0 : 0u, row1 < N ? w1 : 0u, lane);
    }
  }

  // ---- pass B: key tiles -> dK, dV (transposed scores: rows = keys, columns = queries) ----
  for (int jt = warp; jt < NT; jt += nwarps) {
    uint32_t ka[4][4], va[4][4];
    load_a_frags(ka, Ks, jt * 16, lane);
    load_a_frags(va, Vs, jt * 16, lane);
    const int key0 = jt * 16 + g, key1 = key0 + 8;
    float dk[8][4], dv[8][4];
#pragma unroll
    for (int n = 0; n < 8; ++n)
#pragma unroll
      for (int i = 0; i < 4; ++i) dk[n][i] = dv[n][i] = 0.f;
#pragma unroll 1
    for (int it = 0; it < NT; ++it) {
      float s[2][4], dp[2][4];
      mma_rowsT(s, ka, Qs, it * 16, lane);      // S^T tile: (key, query)
      mma_rowsT(dp, va, Gs, it * 16, lane);     // dP^T tile
      float pv[2][4], dsv[2][4];
#pragma unroll
      for (int j = 0; j < 2; ++j)
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          const int q = it * 16 + j * 8 + 2 * t + (i & 1);      // query index = column
          const int key = (i < 2) ? key0 : key1;
          const float p = (key < N) ? __expf(s[j][i] * scale - lse[q]) : 0.f;   // lse[q >= N] = +inf -> 0
          pv[j][i] = p;
          dsv[j][i] = p * (dp[j][i] - Dr[q]);
        }
      uint32_t pf[4], df[4];
      pf[0] = pack2(pv[0][0], pv[0][1]); pf[1] = pack2(pv[0][2], pv[0][3]);
      pf[2] = pack2(pv[1][0], pv[1][1]); pf[3] = pack2(pv[1][2], pv[1][3]);
      df[0] = pack2(dsv[0][0], dsv[0][1]); df[1] = pack2(dsv[0][2], dsv[0][3]);
      df[2] = pack2(dsv[1][0], dsv[1][1]); df[3] = pack2(dsv[1][2], dsv[1][3]);
      mma_rows(dv, pf, Gs, it * 16, lane);
      mma_rows(dk, df, Qs, it * 16, lane);
    }
#pragma unroll
    for (int n = 0; n < 8; ++n) {
      const uint32_t k0 = pack2(dk[n][0] * scale, dk[n][1] * scale), k1 = pack2(dk[n][2] * scale, dk[n][3] * scale);
      const uint32_t v0 = pack2(dv[n][0], dv[n][1]), v1 = pack2(dv[n][2], dv[n][3]);
      if (key0 < N) {
        *reinterpret_cast<uint32_t*>(dq_base + d + (size_t)key0 * rs + n * 8 + 2 * t) = k0;
        *reinterpret_cast<uint32_t*>(dq_base + 2 * d + (size_t)key0 * rs + n * 8 + 2 * t) = v0;
      }
      if (key1 < N) {
        *reinterpret_cast<uint32_t*>(dq_base + d + (size_t)key1 * rs + n * 8 + 2 * t) = k1;
        *reinterpret_cast<uint32_t*>(dq_base + 2 * d + (size_t)key1 * rs + n * 8 + 2 * t) = v1;
      }
      if (dbias != nullptr) {
        tile_colsum(s_db + HD + n * 8 + 2 * t, key0 < N ? k0 : 0u, key1 < N ? k1 : 0u, lane);
        tile_colsum(s_db + 2 * HD + n * 8 + 2 * t, key0 < N ? v0 : 0u, key1 < N ? v1 : 0u, lane);
      }
    }
  }
  if (dbias != nullptr) {     // one global atomic per column per head: qkv bias gradient [3][H][64]
    __syncthreads();
    for (int i = threadIdx.x; i < 3 * HD; i += blockDim.x)
      atomicAdd(dbias + (i / HD) * d + h * HD + (i % HD), s_db[i]);
  }
}

template <int NPAD>
int launch_bwd(const __nv_bfloat16* qkv, const __nv_bfloat16* o, const __nv_bfloat16* dout, const float* lse,
               __nv_bfloat16* dqkv, float* dbias, int B, int N, int H, float scale, cudaStream_t st) {
  const int smem = 4 * NPAD * LDS * 2 + 2 * NPAD * 4 + 3 * HD * 4;
  // warps per CTA chosen so the NPAD/16 row tiles split evenly (13 tiles -> 7 warps x 2 rounds)
  constexpr int NW = NPAD == 208 ? 7 : (NPAD == 256 ? 8 : 4);
  FC_SMEM_OPT_IN((attn_bwd_kernel<NPAD, NW>), smem);
  attn_bwd_kernel<NPAD, NW><<<B * H, NW * 32, smem, st>>>(qkv, o, dout, lse, dqkv, dbias, N, H, scale);
  FC_LAUNCH_CHECK();
  return FC_OK;
}

}  // namespace

extern "C" int fc_attention_bwd(const void* qkv, const void* out, const void* d_out, const float* lse, void* dqkv,
                                float* dbias, int B, int N, int H, int head_dim, int device, void* stream) {
  FC_REQUIRE(head_dim == HD, "fc_attention_bwd: head_dim must be 64 (got %d)", head_dim);
  FC_REQUIRE(B > 0 && N > 0 && N <= 256 && H > 0, "fc_attention_bwd: unsupported shape B=%d N=%d H=%d", B, N, H);
  FcDeviceGuard guard(device);
  const float scale = 0.125f;
  auto q = reinterpret_cast<const __nv_bfloat16*>(qkv);
  auto o = reinterpret_cast<const __nv_bfloat16*>(out);
  auto g = reinterpret_cast<const __nv_bfloat16*>(d_out);
  auto dq = reinterpret_cast<__nv_bfloat16*>(dqkv);
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  if (N <= 48) return launch_bwd<48>(q, o, g, lse, dq, dbias, B, N, H, scale, st);
  if (N <= 64) return launch_bwd<64>(q, o, g, lse, dq, dbias, B, N, H, scale, st);
  if (N <= 208) return launch_bwd<208>(q, o, g, lse, dq, dbias, B, N, H, scale, st);
  return launch_bwd<256>(q, o, g, lse, dq, dbias, B, N, H, scale, st);
}
