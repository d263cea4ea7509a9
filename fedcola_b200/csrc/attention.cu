// tcgen05 fused short-sequence attention, forward and backward (head_dim 64, N <= 256 tokens).
//
// Replaces Attention.forward between the qkv and proj Linears (/root/reference/src/models/mome.py:153-165) —
// q*scale, fp32 QK^T, fp32 softmax, cast, P·V, head merge — and its autograd.
//
// Forward.  Persistent CTAs (one per SM) walk the (sample, head) items; 16 softmax warps + 1 control warp whose
// lane 0 is the TMA producer and the only tcgen05.mma issuer:
//   TMA (3-D tensor maps over qkv [B, N, 3*H*64], 128-byte swizzle) stages Q, K, V of an item in shared memory —
//   token rows >= N are zero-filled by the TMA unit; up to 4 items are in flight;
//   S = Q K^T   : tcgen05.mma  M=128 queries, N=NK (keys padded to 16), K=64      -> TMEM, double-buffered
//   softmax     : TMEM lane = query row; the 4 warps of a lane quarter split the key columns 64 each; pass 1 row
//                 max, pass 2 exp2 + row sum, both straight out of TMEM; partial max/sum go through 6 KB of smem
//                 with one named barrier per tile; the bf16 (un-normalised) P is written back over the S columns
//                 it came from (tcgen05.st);
//   O = P V     : tcgen05.mma  M=128, N=64, K=NK, A = P read from tensor memory, B = V MN-major from smem
//   epilogue    : (one tile later, under the next tile's softmax) O / rowsum -> bf16 -> 32-byte row segments to
//                 global (rows >= N skipped); LSE = max + log(sum) for the backward.
// Backward: see the comment above attn_bwd_tc_kernel.
#include "common.cuh"
#include "sm100.cuh"
#include "../../include/fedcola_b200.h"

#include <cstring>
#include <mutex>

namespace {

using namespace sm100;

constexpr int HD = 64;
constexpr int TILE = 128 * 128;            // bytes of one 128-row x 128-byte swizzled tile (16 KB)
#ifndef FC_ATTN_POLY_EXP2
// which of every 4 column pairs of the forward's exp2 pass take the FMA-pipe polynomial (bit e) instead of MUFU.EX2.
// Measured at B=112 N=197 H=6: 0 -> 32.8 us, 0x2 (25 %) -> 33.5, 0xA (50 %) -> 35.2, 0xE (75 %) -> 36.6: the pass is bound
// by instruction issue and TMEM latency, not by the MUFU rate, so the default is off (profiles/experiments/README.md).
#define FC_ATTN_POLY_EXP2 0x0
#endif
constexpr int kPolyExp2Mask = FC_ATTN_POLY_EXP2;
#ifdef FC_ATTN_PROF
__device__ long long g_attn_bprof[2][32 * 16];      // backward: [0] elementwise warp 0, [1] control lane; 32 chunks x 16 stamps
#define BPROF(who, chunk, slot) do { if (blockIdx.x == 0 && (chunk) < 32) g_attn_bprof[who][(chunk) * 16 + (slot)] = clock64(); } while (0)
__device__ long long g_attn_prof[16 * 12];
#define PROF(slot) do { if (threadIdx.x == 0 && T < 16) prof_s[T * 12 + (slot)] = clock64(); } while (0)
#else
#define PROF(slot) do { } while (0)
#define BPROF(who, chunk, slot) do { } while (0)
#endif
#ifdef FC_ATTN_PROF
constexpr int kMaxDynSmem = 232448 - 2048;
#else
constexpr int kMaxDynSmem = 232448;
#endif
constexpr int kSoftmaxWarps = 16;          // warp&3 = TMEM lane quarter (query rows), warp>>2 = 64-column group
constexpr int kFwdThreads = (kSoftmaxWarps + 1) * 32;   // + one control warp (TMA producer / MMA issuer)

__device__ __forceinline__ uint32_t pack2(float a, float b) {
  __nv_bfloat162 v = __floats2bfloat162_rn(a, b);
  return *reinterpret_cast<uint32_t*>(&v);
}
__device__ __forceinline__ void sts_u4(uint32_t addr, uint32_t a, uint32_t b, uint32_t c, uint32_t d) {
  asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "r"(a), "r"(b), "r"(c), "r"(d) : "memory");
}
__device__ __forceinline__ void st_shared_f32(uint32_t addr, float v) {
  asm volatile("st.shared.f32 [%0], %1;" ::"r"(addr), "f"(v) : "memory");
}
__device__ __forceinline__ float ld_shared_f32(uint32_t addr) {
  float v;
  asm volatile("ld.shared.f32 %0, [%1];" : "=f"(v) : "r"(addr) : "memory");
  return v;
}
__device__ __forceinline__ void st_shared_bf16(uint32_t addr, float v) {
  const __nv_bfloat16 h = __float2bfloat16_rn(v);
  asm volatile("st.shared.b16 [%0], %1;" ::"r"(addr), "h"(*reinterpret_cast<const uint16_t*>(&h)) : "memory");
}
__device__ __forceinline__ float ld_shared_bf16(uint32_t addr) {
  uint16_t h;
  asm volatile("ld.shared.b16 %0, [%1];" : "=h"(h) : "r"(addr) : "memory");
  return __uint_as_float(static_cast<uint32_t>(h) << 16);
}
__device__ __forceinline__ void ld_shared_f32x4(uint32_t addr, float (&v)[4]) {
  asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v[0]), "=f"(v[1]), "=f"(v[2]), "=f"(v[3]) : "r"(addr) : "memory");
}
__device__ __forceinline__ float lg2(float x) {
  float y;
  asm("lg2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
__device__ __forceinline__ float rcp(float x) {
  float y;
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
__device__ __forceinline__ float ex2(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
// exp2 of two scores on the FMA pipes (the forward's exp2 pass is MUFU-bound: 16 results per clock and SM):
// x = v*c - max*c <= 0 clamped to -126; n = round(x) by the 1.5*2^23 trick, f = x - n in [-0.5, 0.5];
// 2^f by a degree-3 minimax polynomial (7.5e-5 relative, P is rounded to bf16 = 2e-3); 2^n goes into the exponent
// field with one shift-add.  FFMA2 / FADD2: two columns per issue slot.
__device__ __forceinline__ void ex2_pair_poly(float v0, float v1, float sc, float nmb, float& p0, float& p1) {
  float x0, x1;
  f2_unpack(f2_fma(f2_pack(v0, v1), f2_all(sc), f2_all(nmb)), x0, x1);
  const uint64_t x = f2_pack(fmaxf(x0, -126.0f), fmaxf(x1, -126.0f));
  const uint64_t t = f2_add(x, f2_all(12582912.0f));
  const uint64_t f = f2_add(x, f2_fma(t, f2_all(-1.0f), f2_all(12582912.0f)));        // x - (t - magic)
  uint64_t q = f2_fma(f, f2_all(0.0551716685295105f), f2_all(0.2426111251115799f));
  q = f2_fma(q, f, f2_all(0.6932609677314758f));
  q = f2_fma(q, f, f2_all(0.9999280571937561f));
  float q0, q1, t0, t1;
  f2_unpack(q, q0, q1);
  f2_unpack(t, t0, t1);
  p0 = __uint_as_float(__float_as_uint(q0) + (__float_as_uint(t0) << 23));
  p1 = __uint_as_float(__float_as_uint(q1) + (__float_as_uint(t1) << 23));
}
__device__ __forceinline__ void softmax_warps_sync() { asm volatile("bar.sync 1, 512;" ::: "memory"); }
// MN-major B operand made of ONE 64-wide block (V: rows = keys = K index, 128-byte rows of 64 head-dim values)
__device__ __forceinline__ uint64_t desc_mn(uint32_t addr) { return umma_smem_desc(addr, 8192, 1024); }
__device__ __forceinline__ uint64_t desc_k(uint32_t addr) { return umma_smem_desc(addr, 16, 1024); }
// The start-address field is the low 14 bits (address >> 4) and shared memory ends below 2^18 bytes, so a descriptor
// for another address is the zero-address descriptor plus (address >> 4): the MMA issue loops below advance their
// operands with one 64-bit add instead of rebuilding the descriptor (the single issuing thread is a bottleneck).
__device__ __forceinline__ uint64_t desc_at(uint64_t zero_addr_desc, uint32_t addr) { return zero_addr_desc + (addr >> 4); }

// Client groups (FC_ATTN_MAX_GROUPS): one launch walks the (sample, head) items of several clients' qkv tensors of the
// same shape — item = (group, sample, head) — so the persistent kernel's fixed costs are paid once per client group.
constexpr int MAXG = FC_ATTN_MAX_GROUPS;
struct FwdMaps {
  CUtensorMap qkv_a, qkv_b;                  // box rows RA (first 128-token tile) / RB (remainder tile)
  CUtensorMap out_st;                        // output store: box = 64 columns x RA rows (rows past N are clipped)
};
struct FwdGroups {
  FwdMaps maps[MAXG];
  __nv_bfloat16* out[MAXG];
  float* lse[MAXG];
};

// Persistent: CTA c handles (sample, head) items c, c+grid, ...  Warp 16 is the control warp: its lane 0 prefetches
// the next items' Q/K/V by TMA (nbuf smem buffers) and issues every tcgen05.mma; warps 0-15 do the softmax and the
// output.  S accumulators are double-buffered in TMEM (columns 0 / 256) and O overlays its own consumed S, so
// QK^T of tile T+1 and P·V of tile T run under the softmax warps' work on the neighbouring tiles.
template <bool STAGED>                       // STAGED: O leaves through staging tiles + TMA stores (own code, own registers)
__global__ void __launch_bounds__(kFwdThreads, 1)   // 17 warps = 5 on one SM sub-partition: 96 registers per thread at most
attn_fwd_tc_kernel(const __grid_constant__ FwdGroups G, int n_items, int items_per_group, int N, int H, int nbuf,
                   float scale_log2e) {
  constexpr bool staged = STAGED;
  extern __shared__ __align__(1024) uint8_t smem[];   // swizzled tiles need 1024-byte alignment (checked below)
#ifdef FC_ATTN_PROF
  __shared__ long long prof_s[16 * 12];
#endif
  if (smem_u32(smem) & 1023) __trap();
  const int RA = N > 128 ? 128 : ((N + 15) & ~15);
  const int RB = N > 128 ? ((N - 128 + 15) & ~15) : 0;
  const int NK = RA + RB;                     // keys padded to the UMMA N granularity; rows >= N are TMA zero fill
  const int q_tiles = RB ? 2 : 1;
  const int op_bytes = NK * 128;              // one operand (Q, K or V) of one item
  const int buf_bytes = 3 * op_bytes;
  // reduction scratch, double-buffered by tile parity: row-sum partials float [2][4][128], then row-max partials
  // bf16 [2][4][128] (softmax is shift-invariant: a rounded max only has to be the SAME for the whole row)
  // staged: two tiles between the operand buffers and the scratch, through which O leaves as TMA stores
  uint8_t* stage_ptr = smem + nbuf * buf_bytes;
  uint8_t* red_base = stage_ptr + (staged ? 2 * TILE : 0);
  uint64_t* tma_bar = reinterpret_cast<uint64_t*>(red_base + 6144);    // [4] item operands landed
  uint64_t* s_full = tma_bar + 4;                                      // [2] S buffer written by the tensor core
  uint64_t* o_full = s_full + 2;                                       //     O written (and P, V no longer read)
  uint64_t* p_full = o_full + 1;                                       //     P tile written by the 16 softmax warps
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(p_full + 1);
  const int d = H * HD;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

  if (threadIdx.x == 0) {
    for (int i = 0; i < 4; ++i) mbar_init(&tma_bar[i], 1);
    mbar_init(&s_full[0], 1);
    mbar_init(&s_full[1], 1);
    mbar_init(o_full, 1);
    mbar_init(p_full, kSoftmaxWarps);
    fence_barrier_init();
  }
  if (warp == kSoftmaxWarps) {
    tmem_alloc(tmem_slot, 512);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *tmem_slot;
  // TMEM columns: NK <= 208: S0 [0,208) S1 [208,416) O [416,480); larger NK: one S buffer [0,256), O [256,320)
  const int depth = NK <= 208 ? 2 : 1;
  const uint32_t s_stride = 208, o_col = NK <= 208 ? 416 : 256;
  const int first = blockIdx.x, stride = gridDim.x;
  const int n_my = first < n_items ? (n_items - first + stride - 1) / stride : 0;

  if (warp == kSoftmaxWarps) {
    // ================= control warp: TMA producer + MMA issuer =================
    // warp-uniform control flow, elect-predicated TMA / MMA / commit (see the backward's control warp)
    if (n_my > 0) {
      const bool leader = elect_one();
      for (int g = 0; g * items_per_group < n_items; ++g) {
        if (leader) prefetch_tmap(&G.maps[g].qkv_a);
        if (leader) prefetch_tmap(&G.maps[g].qkv_b);
      }
      const uint32_t idesc_s = umma_idesc_bf16(128, NK, 0, 0);
      const uint32_t idesc_o = umma_idesc_bf16(128, HD, 0, 1);
      auto buf_addr = [&](int k) { return smem_u32(smem) + (k % nbuf) * buf_bytes; };
      auto issue_load = [&](int k) {
        const int item = first + k * stride;
        const int grp = item / items_per_group, rem = item - grp * items_per_group;
        const int b = rem / H, h = rem % H;
        const FwdMaps& maps = G.maps[grp];
        uint8_t* base = smem + (k % nbuf) * buf_bytes;
        uint64_t* bar = &tma_bar[k % nbuf];
        if (leader) mbar_arrive_expect_tx(bar, buf_bytes);
        for (int op = 0; op < 3; ++op) {      // Q, K, V column blocks of this head
          if (leader) tma_load_3d(base + op * op_bytes, &maps.qkv_a, bar, op * d + h * HD, 0, b);
          if (RB && leader) tma_load_3d(base + op * op_bytes + RA * 128, &maps.qkv_b, bar, op * d + h * HD, 128, b);
        }
      };
      const uint64_t dk0 = desc_k(0), dmn0 = desc_mn(0);
      auto issue_s = [&](int k, int qt, int sbuf) {   // S[sbuf] = Q_tile K^T
        if (qt == 0) mbar_wait(&tma_bar[k % nbuf], (k / nbuf) & 1);
        tc_fence_after();
        const uint64_t q = desc_at(dk0, buf_addr(k) + qt * TILE), kk = desc_at(dk0, buf_addr(k) + op_bytes);
        const uint32_t ts = tmem + sbuf * s_stride;
#pragma unroll
        for (int j = 0; j < 4; ++j) if (leader) umma_bf16(ts, q + 2 * j, kk + 2 * j, idesc_s, j > 0);   // 32 bytes per k-step
        if (leader) umma_commit(&s_full[sbuf]);
      };
      // Issue order is a small state machine: loads run up to nbuf items ahead (a buffer is reusable once the last
      // P·V of its item completed), QK^T runs up to `depth` tiles ahead of the softmax (an S buffer is reusable
      // once the softmax warps consumed it, i.e. p_full of that tile), P·V follows p_full.
      const int total = n_my * q_tiles;
      int next_load = 0, next_s = 0, done_items = 0;
      auto pump = [&](int T) {
        while (next_load < n_my && next_load < done_items + nbuf) issue_load(next_load++);
        while (next_s < total && next_s < T + depth && next_s / q_tiles < next_load) {
          issue_s(next_s / q_tiles, next_s % q_tiles, next_s % depth);
          ++next_s;
        }
      };
      pump(0);
      for (int T = 0; T < total; ++T) {
        const int k = T / q_tiles;
        while (!mbar_try_wait(p_full, T & 1)) __nanosleep(100);   // P_T in smem, S_T consumed, O_{T-1} read
        tc_fence_after();
        uint64_t vv = desc_at(dmn0, buf_addr(k) + 2 * op_bytes);
        const int ksteps = NK >> 4;
        // P (bf16 pairs) sits in the S buffer: the 16 keys of k-step ks occupy 8 columns at 64*(ks/4) + 32*((ks/2)&1) +
        // 8*(ks&1) — each softmax warp packed its 32-score slabs in place (slab at +32*half, packed into its first 16)
        const uint32_t tP = tmem + (depth == 2 ? (T & 1) : 0) * s_stride;
        for (int ks = 0; ks < ksteps; ++ks, vv += 128)     // 16 key rows of V = 2048 bytes
          if (leader) umma_bf16_ts(tmem + o_col, tP + 64 * (ks >> 2) + 32 * ((ks >> 1) & 1) + 8 * (ks & 1), vv, idesc_o, ks > 0);
        if (leader) umma_commit(o_full);
        pump(T + 1);
        if ((T + 1) % q_tiles == 0) {
          mbar_wait(o_full, T & 1);           // every MMA reading buffer k % nbuf has completed
          ++done_items;
          pump(T + 1);
        }
      }
    }
  } else {
    // ================= softmax warps =================
    const int quarter = warp & 3, grp = warp >> 2;
    const int row = quarter * 32 + lane;      // query row inside the tile == TMEM lane
    const uint32_t lane_off = static_cast<uint32_t>(quarter * 32) << 16;
    const int c0 = grp * 64;
    const int nc = min(max(NK - c0, 0), 64);  // columns of this warp (multiple of 16)
    const int valid = min(nc, N - c0);        // ... of which real keys
    const uint32_t red_sum = smem_u32(red_base) + (grp * 128 + row) * 4, red_sum_rd = smem_u32(red_base) + row * 4;
    const uint32_t red_max = smem_u32(red_base) + 4096 + (grp * 128 + row) * 2;
    const uint32_t red_max_rd = smem_u32(red_base) + 4096 + row * 2;
    // one 32-wide (or trailing 16-wide) slab of this thread's score row; returns the number of columns loaded
    auto load_slab = [&](uint32_t tS, int half, float (&v)[32]) -> int {
      const int n = min(nc - 32 * half, 32);
      if (n == 32) tmem_ld_32x32(tS + lane_off + c0 + 32 * half, v);
      else if (n == 16) tmem_ld_32x16(tS + lane_off + c0 + 32 * half, *reinterpret_cast<float(*)[16]>(&v[0]));
      tmem_ld_wait();
      return n;
    };
    // O of a finished tile (columns [16*grp, +16) of this thread's row), scaled by that tile's 1/rowsum.  Loaded
    // before this tile's P is published (the next P·V overwrites O), stored after it (a proxy fence behind
    // outstanding global stores would wait for them).
    auto load_o = [&](int T, float inv, uint32_t (&o)[8]) {
      mbar_wait(o_full, T & 1);
      tc_fence_after();
      float f[16];
      tmem_ld_32x16(tmem + o_col + lane_off + grp * 16, f);
      tmem_ld_wait();
#pragma unroll
      for (int i = 0; i < 8; ++i) o[i] = pack2(f[2 * i] * inv, f[2 * i + 1] * inv);
    };
    // staged output: the thread's 32 bytes go into the swizzled staging tile of the tile's parity; thread 0 hands the
    // tile to the TMA engine after the next all-warp barrier (per-thread 32-byte global stores cost the LSU one
    // request per sector: ~2 k cycles per tile with every softmax warp waiting)
    const uint32_t sStage = smem_u32(stage_ptr);
    auto stage_o = [&](const uint32_t (&o)[8], int T_of) {
      const uint32_t rsw = (sStage + (T_of & 1) * TILE + row * 128) | ((row & 7) << 4);
      sts_u4(rsw ^ ((2 * grp) << 4), o[0], o[1], o[2], o[3]);
      sts_u4(rsw ^ ((2 * grp + 1) << 4), o[4], o[5], o[6], o[7]);
      fence_proxy_async();
    };
    auto issue_store = [&](int T_of, int b, int h, int qt, int cg) {      // one thread
      tma_store_3d(&G.maps[cg].out_st, stage_ptr + (T_of & 1) * TILE, h * HD, qt * 128, b);
      tma_store_commit();
    };
    auto store_o = [&](const uint32_t (&o)[8], int bh_row0, int h, int qt, int cg) {
      const int qrow = qt * 128 + row;
      if (qrow < N) {
        uint4* dst = reinterpret_cast<uint4*>(G.out[cg] + (static_cast<size_t>(bh_row0) + qrow) * d + h * HD + grp * 16);
        dst[0] = make_uint4(o[0], o[1], o[2], o[3]);
        dst[1] = make_uint4(o[4], o[5], o[6], o[7]);
      }
    };
    // keys >= N of a slab (columns at or beyond `nvalid`) -> -inf; only the slab holding the padding boundary
    auto mask_slab = [&](float (&v)[32], int nvalid) {
#pragma unroll
      for (int i = 0; i < 32; ++i) v[i] = i < nvalid ? v[i] : -INFINITY;
    };
    const float sc = scale_log2e;
    // row sums of the previous tile -> 1/sum for its O rows, LSE for the backward
    auto finish_sums = [&](int T, float mb, int bh, int qt, int cg) -> float {     // bh: item index inside its client group
      const uint32_t rd = red_sum_rd + (T & 1) * 2048;
      const float sum = (ld_shared_f32(rd) + ld_shared_f32(rd + 512)) + (ld_shared_f32(rd + 1024) + ld_shared_f32(rd + 1536));
      const int qrow = qt * 128 + row;
      float* lse_out = G.lse[cg];
      if (grp == 0 && lse_out != nullptr && qrow < N)
        lse_out[static_cast<size_t>(bh) * N + qrow] = 0.69314718056f * (mb + lg2(sum));
      return rcp(sum);
    };
    int T = 0, prev_row0 = 0, prev_h = 0, prev_qt = 0, prev_bh = 0, prev_cg = 0, prev_b = 0;
    int prev2_b = 0, prev2_h = 0, prev2_qt = 0, prev2_cg = 0;         // tile T-2: its O is staged, not yet handed over
    float prev_mb = 0.f;
    const unsigned long long h_magic = ((1ull << 32) + H - 1) / H;   // item / H == (item * magic) >> 32 for item < 2^16
    for (int k = 0, item = first; k < n_my; ++k, item += stride) {
      const int cg = item / items_per_group, rem = item - cg * items_per_group;
      const int b = static_cast<int>((static_cast<unsigned long long>(rem) * h_magic) >> 32), h = rem - b * H;
      for (int qt = 0; qt < q_tiles; ++qt, ++T) {
        const int sb = depth == 2 ? (T & 1) : 0;
        const uint32_t tS = tmem + sb * s_stride;
        const uint32_t par = (T & 1) * 1024;  // reduction scratch is double-buffered by tile parity: one barrier/tile
        PROF(0);
        mbar_wait(&s_full[sb], (depth == 2 ? (T >> 1) : T) & 1);
        tc_fence_after();
        PROF(1);
        // ---- pass 1: row max over this warp's columns (keys >= N masked) ----
        float m0 = -INFINITY, m1 = -INFINITY;
#pragma unroll
        for (int half = 0; half < 2; ++half) {
          if (32 * half < nc) {                // warp-uniform
            float v[32];
            const int n = load_slab(tS, half, v);
            if (32 * half + n > valid) mask_slab(v, valid - 32 * half);
#pragma unroll
            for (int i = 0; i < 32; i += 2) {
              if (i < n) {
                m0 = fmaxf(m0, v[i]);
                m1 = fmaxf(m1, v[i + 1]);
              }
            }
          }
        }
        st_shared_bf16(red_max + par, fmaxf(m0, m1));
        if (staged && threadIdx.x == 0) tma_store_wait_read();       // the staging tile written below is free again
        softmax_warps_sync();                 // also: every warp has published the previous tile's row sums
        if (staged && T >= 2 && threadIdx.x == 0) issue_store(T - 2, prev2_b, prev2_h, prev2_qt, prev2_cg);
        PROF(3);
        const float mx = fmaxf(fmaxf(ld_shared_bf16(red_max_rd + par), ld_shared_bf16(red_max_rd + par + 256)),
                               fmaxf(ld_shared_bf16(red_max_rd + par + 512), ld_shared_bf16(red_max_rd + par + 768)));
        const float mb = mx * sc, nmb = -mb;
        // the previous tile's P·V ran under the work above: fetch its O now (this also frees the P tile)
        uint32_t o[8];
        if (T > 0) load_o(T - 1, finish_sums(T - 1, prev_mb, prev_bh, prev_qt, prev_cg), o);
        PROF(4);
        // ---- pass 2: p = exp2(s*c - max*c) -> bf16 -> swizzled K-major P tile (column block = grp); row sum ----
        // P is left un-normalised (0 < p <= 1); the row of O is scaled by 1/sum in its epilogue.
        float s0 = 0.f, s1 = 0.f;
#pragma unroll
        for (int half = 0; half < 2; ++half) {
          if (32 * half < nc) {
            float v[32];
            const int n = load_slab(tS, half, v);
            if (32 * half + n > valid) mask_slab(v, valid - 32 * half);   // exp2(-inf) = 0
#pragma unroll
            for (int j = 0; j < 4; ++j) {
              if (8 * j < n) {
                uint32_t pk[4];
#pragma unroll
                for (int e = 0; e < 4; ++e) {
                  const int i = 8 * j + 2 * e;
                  float p0, p1;
                  if (kPolyExp2Mask & (1 << e)) {            // this pair on the FMA pipes, the others on the MUFU
                    ex2_pair_poly(v[i], v[i + 1], sc, nmb, p0, p1);
                  } else {
                    p0 = ex2(fmaf(v[i], sc, nmb));
                    p1 = ex2(fmaf(v[i + 1], sc, nmb));
                  }
                  s0 += p0;
                  s1 += p1;
                  pk[e] = pack2(p0, p1);
                }
                // 8 keys -> 4 packed columns, written over this thread's own (already consumed) scores
                tmem_st_32x4(tS + lane_off + c0 + 32 * half + 4 * j, pk[0], pk[1], pk[2], pk[3]);
              }
            }
          }
        }
        st_shared_f32(red_sum + 2 * par, s0 + s1);
        tmem_st_wait();                       // P is in tensor memory (A operand of P·V)
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(p_full);
        PROF(5);
        if (T > 0) {
          if (staged) stage_o(o, T - 1);
          else store_o(o, prev_row0, prev_h, prev_qt, prev_cg);
        }
        prev2_b = prev_b;
        prev2_h = prev_h;
        prev2_qt = prev_qt;
        prev2_cg = prev_cg;
        prev_b = b;
        prev_mb = mb;
        prev_row0 = b * N;
        prev_bh = rem;
        prev_h = h;
        prev_qt = qt;
        prev_cg = cg;
        PROF(6);
      }
    }
    if (T > 0) {
      if (staged && threadIdx.x == 0) tma_store_wait_read();
      softmax_warps_sync();                   // last tile's row sums
      if (staged && T >= 2 && threadIdx.x == 0) issue_store(T - 2, prev2_b, prev2_h, prev2_qt, prev2_cg);
      uint32_t o[8];
      load_o(T - 1, finish_sums(T - 1, prev_mb, prev_bh, prev_qt, prev_cg), o);
      if (staged) {
        stage_o(o, T - 1);
        softmax_warps_sync();
        if (threadIdx.x == 0) {
          issue_store(T - 1, prev_b, prev_h, prev_qt, prev_cg);
          tma_store_wait_all();
        }
      } else {
        store_o(o, prev_row0, prev_h, prev_qt, prev_cg);
      }
    }
  }
  tc_fence_before();
  __syncthreads();
#ifdef FC_ATTN_PROF
  if (blockIdx.x == 0 && threadIdx.x < 16 * 12) g_attn_prof[threadIdx.x] = prof_s[threadIdx.x];
#endif
  if (warp == kSoftmaxWarps) {
    tc_fence_after();
    tmem_dealloc(tmem, 512);
  }
}

// =====================================================================================================
// Backward.  Per (sample, head) item, in the S^T orientation (TMEM lane = key row, columns = queries):
//   for each 128-key tile kt, for each 64-query chunk qc:
//     S^T  = K[kt] Q[qc]^T, dP^T = V[kt] dO[qc]^T                      (tcgen05, SS)   -> TMEM chunk buffer (x2)
//     P^T  = exp2(S^T*c - lse2[q]),  dS^T = P^T ∘ (dP^T - D[q])         (16 warps; lane = key, 16 columns each)
//            P^T (bf16 pairs) is written back over the thread's own S^T columns, dS^T to a swizzled smem tile
//     dV[kt] += P^T dO[qc]                                             (tcgen05, A from TMEM, B = dO MN-major)
//     per pair of chunks (128 queries):
//     dK[kt] += dS^T Q[pair]                                           (tcgen05, SS, B = Q MN-major)
//     dQ[pair] += dS K[kt]                                             (tcgen05, SS, A = dS^T tile read MN-major)
//   dV/dK leave TMEM after each key tile, dQ (2 x 64 columns) after the item; 1/sqrt(d) is applied there.
// TMEM columns: chunk buffers 2 x (64 S^T + 64 dP^T) = [0,256), dV [256,320), dK [320,384), dQ [384,512).
// D[q] = <dO[q], O[q]> and lse2[q] = lse[q]*log2(e) are computed from global memory one item ahead.
// The qkv bias gradient is taken in the accumulator epilogues: the Q third by one warp reduction per item, the V
// third from an all-ones row of P^T (sum_k dV = sum_q dO), the K third is identically zero.
constexpr int kVecWarps = 2;                // warps that compute lse2 / D of the next item while this one runs
constexpr int kBwdThreads = (kSoftmaxWarps + 1 + kVecWarps) * 32;
constexpr int kRing = 4;                   // dS^T tiles (64 queries each): two pairs in flight

struct BwdMaps {
  CUtensorMap qkv_a, qkv_b, do_a, do_b;      // box rows RA / RB
  CUtensorMap dqkv_st;                       // gradient store: box = 64 columns x RA rows (rows past N are clipped)
};
struct BwdGroups {
  BwdMaps maps[MAXG];
  const __nv_bfloat16* o[MAXG];
  const __nv_bfloat16* d_o[MAXG];
  const float* lse[MAXG];
  __nv_bfloat16* dqkv[MAXG];
  float* dbias[MAXG];
};

template <bool ONE_CHUNK>                   // ONE_CHUNK: N <= 64, an item is a single chunk (own flush path, own code)
__global__ void __launch_bounds__(kBwdThreads, 1)   // (19 warps: 96 registers is what the file grants in units of 512 per warp)
attn_bwd_tc_kernel(const __grid_constant__ BwdGroups G, int n_items, int items_per_group, int N, int H, float scale,
                   int staged) {
  extern __shared__ __align__(1024) uint8_t smem[];
  if (smem_u32(smem) & 1023) __trap();
  const int RA = N > 128 ? 128 : ((N + 15) & ~15);
  const int RB = N > 128 ? ((N - 128 + 15) & ~15) : 0;
  const int NK = RA + RB;                     // padded tokens (rows >= N are TMA zero fill)
  const int nkt = RB ? 2 : 1;                 // key tiles
  const int nqc = (NK + 63) >> 6;             // query chunks per key tile
  const int op_bytes = NK * 128;
  // One key tile (N <= 128): the operands are small, so there are TWO operand sets and item k+1 loads while item k
  // computes.  Two key tiles: one set, reloaded piecewise as its parts die (see the control warp).
  const int nsets = nkt == 1 ? 2 : 1;
  // smem: nsets x (Q | K | V | dO) | dS^T ring | lse2,D vectors [2 items][2][256] | barriers
  const uint32_t sQ = smem_u32(smem);
  const uint32_t sRing = sQ + nsets * 4 * op_bytes;
  // staged: two more tiles through which dV/dK and dQ leave as TMA stores (when they fit, see the host side)
  uint8_t* stage_ptr = smem + nsets * 4 * op_bytes + kRing * TILE;
  const uint32_t sStage = sRing + kRing * TILE;
  uint8_t* vec_base = stage_ptr + staged * TILE;        // staged = number of staging tiles: 0, 2 or 3
  uint64_t* tma_bar = reinterpret_cast<uint64_t*>(vec_base + 4096);   // [3] operand pieces / operand sets
  uint64_t* sdp_full = tma_bar + 3;           // [2] S^T/dP^T chunk buffer written
  uint64_t* e_done = sdp_full + 2;            // [2] chunk consumed: P^T in TMEM, dS^T in smem          (16 arrivals)
  uint64_t* pair_done = e_done + 2;           // [2] dK/dQ (and every earlier MMA) of a pair completed
  uint64_t* acc_free = pair_done + 2;         //     dV/dK of a key tile read out of TMEM                (16 arrivals)
  uint64_t* dq_free = acc_free + 1;           //     dQ of an item read out of TMEM                      (16 arrivals)
  uint64_t* vec_full = dq_free + 1;           // [2] lse2 / D vectors of an item written                 (kVecWarps arrivals)
  uint64_t* vec_free = vec_full + 2;          // [2] ... and no longer read by the elementwise warps     (16 arrivals)
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(vec_free + 2);
  const int d = H * HD, d3 = 3 * d;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

  if (threadIdx.x == 0) {
    mbar_init(&tma_bar[0], 1);
    mbar_init(&tma_bar[1], 1);
    mbar_init(&tma_bar[2], 1);
    mbar_init(&sdp_full[0], 1);
    mbar_init(&sdp_full[1], 1);
    mbar_init(&e_done[0], kSoftmaxWarps);
    mbar_init(&e_done[1], kSoftmaxWarps);
    mbar_init(&pair_done[0], 1);
    mbar_init(&pair_done[1], 1);
    mbar_init(acc_free, kSoftmaxWarps);
    mbar_init(dq_free, kSoftmaxWarps);
    for (int i = 0; i < 2; ++i) {
      mbar_init(&vec_full[i], kVecWarps);
      mbar_init(&vec_free[i], kSoftmaxWarps);
    }
    fence_barrier_init();
  }
  if (warp == kSoftmaxWarps) {
    tmem_alloc(tmem_slot, 512);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *tmem_slot;
  const uint32_t tDV = tmem + 256, tDK = tmem + 320, tDQ = tmem + 384;
  // a CTA takes a contiguous run of items: consecutive items are the heads of one sample (the same token rows of
  // O / dO / qkv, the same pages), which keeps the per-item global reads of the elementwise warps short
  const int first = static_cast<int>(static_cast<long long>(blockIdx.x) * n_items / gridDim.x), stride = 1;
  const int n_my = static_cast<int>(static_cast<long long>(blockIdx.x + 1) * n_items / gridDim.x) - first;
  const unsigned long long h_magic = ((1ull << 32) + H - 1) / H;
  // item -> (client group, sample, head)
  auto item_bh = [&](int item, int& cg, int& b, int& h) {
    cg = item / items_per_group;
    const int rem = item - cg * items_per_group;
    b = static_cast<int>((static_cast<unsigned long long>(rem) * h_magic) >> 32);
    h = rem - b * H;
  };
  const bool has_dbias = G.dbias[0] != nullptr;       // (all groups agree: checked by the host)

  if (warp == kSoftmaxWarps) {
    // ================= control warp: TMA producer + MMA issuer =================
    // The whole warp runs the control flow (uniform values -> the descriptors live in uniform registers) and only
    // the TMA / MMA / commit instructions are predicated on the elected lane.  Issuing from inside an `if (lane == 0)`
    // branch makes the compiler wrap every tcgen05.mma in an ELECT / BRA.U.ANY loop: >= 46-77 cycles per
    // instruction against the 32 (A in TMEM) - 48 (A in shared memory) cycles an M=128, N=64 MMA needs
    // (tools/mma_probe.cu) — the backward issues ~140 of them per item.
    if (n_my > 0) {
      const bool leader = elect_one();
      for (int g = 0; g * items_per_group < n_items; ++g) {
        if (leader) {
          prefetch_tmap(&G.maps[g].qkv_a);
          prefetch_tmap(&G.maps[g].qkv_b);
          prefetch_tmap(&G.maps[g].do_a);
          prefetch_tmap(&G.maps[g].do_b);
        }
      }
      const uint32_t idesc_dv = umma_idesc_bf16(128, HD, 0, 1);    // A K-major (TMEM / dS^T tile), B MN-major
      const uint32_t idesc_dq = umma_idesc_bf16(128, HD, 1, 1);    // A = dS^T tile read MN-major, B MN-major
      // Operand pieces, each with its own barrier, so that the next item's operands arrive while this item computes:
      //   one key tile : piece 0 = the whole item, into operand set (k & 1)                      [bar = set]
      //   two key tiles: piece 0 = K,V rows [0,128)   dead once the last pair of key tile 0 has completed
      //                  piece 1 = Q,dO rows [0,128)  dead once the first pair of key tile 1 has completed
      //                  piece 2 = rows [128,NK) of all four: dead with the item's last pair; first needed by chunk 2
      auto load_piece = [&](int item, int piece, uint64_t* bar, uint8_t* base) {
        int cg, b, h;
        item_bh(item, cg, b, h);
        const BwdMaps& maps = G.maps[cg];
        if (leader) mbar_arrive_expect_tx(bar, nkt == 1 ? 4 * op_bytes : (piece == 2 ? 4 * RB * 128 : 2 * RA * 128));
        for (int op = 0; op < 4; ++op) {      // Q, K, V (columns of qkv), dO
          const bool kv = op == 1 || op == 2;
          if (nkt == 2 && ((piece == 0 && !kv) || (piece == 1 && kv))) continue;
          const int col = (op < 3 ? op * d : 0) + h * HD;
          if (!leader) continue;
          if (nkt == 1 || piece < 2) tma_load_3d(base + op * op_bytes, op < 3 ? &maps.qkv_a : &maps.do_a, bar, col, 0, b);
          else tma_load_3d(base + op * op_bytes + RA * 128, op < 3 ? &maps.qkv_b : &maps.do_b, bar, col, 128, b);
        }
      };
      auto prefetch_item = [&](int item) {    // into L2 only
        int cg, b, h;
        item_bh(item, cg, b, h);
        const BwdMaps& maps = G.maps[cg];
        for (int op = 0; op < 4; ++op) {
          const int col = (op < 3 ? op * d : 0) + h * HD;
          if (!leader) continue;
          tma_prefetch_3d(op < 3 ? &maps.qkv_a : &maps.do_a, col, 0, b);
          if (RB) tma_prefetch_3d(op < 3 ? &maps.qkv_b : &maps.do_b, col, 128, b);
        }
      };
      int cc = 0, pair = 0, tile_ctr = 0;     // global chunk / pair / key-tile counters
      // A tcgen05.mma of this size costs >= ~48 cycles whatever N is (tools/mma_probe.cu), and every instruction the
      // issuing lane executes between two MMAs adds to that: no divisions, descriptors by addition, unrolled groups.
      const uint64_t dk0 = desc_k(0), dmn0 = desc_mn(0), dq0 = umma_smem_desc(0, TILE, 1024);
      // operand set in use: K-major descriptors (S^T / dP^T operands) and MN-major ones (B of dV, dK, dQ)
      uint64_t dK = 0, dQ = 0, dV = 0, dDO = 0, mQ = 0, mK = 0, mDO = 0;
      auto bind_ahead = [&](int set) {        // operands of S^T / dP^T (these run up to one item ahead)
        const uint32_t q = sQ + set * 4 * op_bytes;
        dQ = desc_at(dk0, q);
        dK = desc_at(dk0, q + op_bytes);
        dV = desc_at(dk0, q + 2 * op_bytes);
        dDO = desc_at(dk0, q + 3 * op_bytes);
      };
      auto bind_cur = [&](int set) {          // B operands of dV / dK / dQ of the item being consumed
        const uint32_t q = sQ + set * 4 * op_bytes;
        mQ = desc_at(dmn0, q);
        mK = desc_at(dmn0, q + op_bytes);
        mDO = desc_at(dmn0, q + 3 * op_bytes);
      };
      bind_ahead(0);
      bind_cur(0);
      const int w_last = NK - 64 * (nqc - 1);                           // queries in the last chunk (16..64)
      const uint32_t idesc_full = umma_idesc_bf16(128, 64, 0, 0), idesc_last = umma_idesc_bf16(128, w_last, 0, 0);
      const int wp16_last = (NK - 128 * ((nqc - 1) >> 1)) >> 4;         // 16-query steps of the last pair (1..8)
      int a_kt = 0, a_qc = 0;                 // the next chunk issue_sdp will compute (cursor inside the item)
      auto issue_sdp = [&](int gc) {          // -> chunk buffer gc & 1
        const uint32_t idesc = a_qc == nqc - 1 ? idesc_last : idesc_full;
        const uint32_t tb = tmem + (gc & 1) * 128;
        const uint64_t ak = dK + a_kt * (TILE >> 4), av = dV + a_kt * (TILE >> 4);
        const uint64_t bq = dQ + a_qc * 512, bd = dDO + a_qc * 512;       // 64 rows x 128 B = 512 x 16 B
#pragma unroll
        for (int j = 0; j < 4; ++j) if (leader) umma_bf16(tb, ak + 2 * j, bq + 2 * j, idesc, j > 0);
#pragma unroll
        for (int j = 0; j < 4; ++j) if (leader) umma_bf16(tb + 64, av + 2 * j, bd + 2 * j, idesc, j > 0);
        if (leader) umma_commit(&sdp_full[gc & 1]);
        if (++a_qc == nqc) {
          a_qc = 0;
          ++a_kt;
        }
      };
      const int chunks = nkt * nqc;
      // the first two chunks of item k: what they read (the item's set, or pieces 0 and 1) has been requested earlier
      auto start_item = [&](int k) {
        if (nkt == 1) {
          mbar_wait(&tma_bar[k & 1], (k >> 1) & 1);
          bind_ahead(k & 1);
        } else {
          mbar_wait(&tma_bar[0], k & 1);
          mbar_wait(&tma_bar[1], k & 1);
        }
        tc_fence_after();
        a_kt = 0;
        a_qc = 0;
        issue_sdp(cc);
        if (chunks > 1) issue_sdp(cc + 1);
      };
      // one key tile (two operand sets): S^T / dP^T of item k+1 are issued while item k is still being consumed —
      // chunk c of item k+1 as soon as the chunk buffer it needs is free
      auto open_next = [&](int k1) {
        mbar_wait(&tma_bar[k1 & 1], (k1 >> 1) & 1);
        tc_fence_after();
        bind_ahead(k1 & 1);
        a_kt = 0;
        a_qc = 0;
      };
      if (n_my > 0) {
        if (nkt == 1) {
          load_piece(first, 0, &tma_bar[0], smem);
          if (n_my > 1) load_piece(first + stride, 0, &tma_bar[1], smem + 4 * op_bytes);
        } else {
          for (int piece = 0; piece < 3; ++piece) load_piece(first, piece, &tma_bar[piece], smem);
          if (n_my > 1) prefetch_item(first + stride);
        }
        start_item(0);
      }
      for (int k = 0; k < n_my; ++k) {
        const int next_item = first + (k + 1) * stride;
        int lc = 0;
        if (nkt == 1) {
          bind_cur(k & 1);
          if (chunks == 1 && k + 1 < n_my) {  // the other chunk buffer is free: item k+1 starts before item k is consumed
            open_next(k + 1);
            issue_sdp(cc + 1);
          }
        }
        for (int kt = 0; kt < nkt; ++kt) {
          const int ksteps = (kt == 0 ? RA : RB) >> 4;
          for (int qc = 0; qc < nqc; ++qc, ++lc, ++cc) {
            const bool last_qc = qc == nqc - 1;
            const int w16 = last_qc ? w_last >> 4 : 4;                  // 16-query steps in this chunk
            if (leader) BPROF(1, cc, 0);
            mbar_wait(&e_done[cc & 1], (cc >> 1) & 1);
            tc_fence_after();
            if (leader) BPROF(1, cc, 1);
            if (qc == 0 && tile_ctr > 0) mbar_wait(acc_free, (tile_ctr - 1) & 1);   // dV/dK of the previous tile read
            if (leader) BPROF(1, cc, 2);
            // dV[kt] += P^T dO[chunk]   (A: packed P^T in this chunk's S^T columns, 8 columns per 16 queries)
            const uint32_t tb = tmem + (cc & 1) * 128;
            {
              const uint64_t b = mDO + qc * 512;
#pragma unroll
              for (int j = 0; j < 4; ++j)
                if (leader && j < w16) umma_bf16_ts(tDV, tb + 16 * j, b + 128 * j, idesc_dv, (qc | j) != 0);
            }
            // the chunk two ahead reuses this chunk's buffer: ordered behind dV above in the MMA pipe, and issued
            // before the pair's dK/dQ so the elementwise warps do not wait for it
            if (lc + 2 < chunks) {
              if (nkt == 2 && lc == 0) {      // chunk 2 is the first to read rows >= 128 (piece 2)
                mbar_wait(&tma_bar[2], k & 1);
                tc_fence_after();
              }
              issue_sdp(cc + 2);
            } else if (nkt == 1 && chunks == 2 && k + 1 < n_my) {
              if (lc == 0) open_next(k + 1);
              issue_sdp(cc + 2);              // chunk lc of item k+1 into the buffer this chunk has just vacated
            }
            // two key tiles: parts of the operand set are dead from here on -> the next item's copy of them.  The
            // pair waited for is the most recently committed one (a parity wait must not fall two phases behind).
            // L2 prefetch of the item after next: early in the item, away from the item boundary where the gradient
            // stores and the piece loads queue up in the TMA unit
            if (nkt == 2 && lc == 2 && k + 2 < n_my) prefetch_item(first + (k + 2) * stride);
            if (nkt == 2 && kt == 1 && (qc == 0 || qc == 2) && k + 1 < n_my) {
              mbar_wait(&pair_done[(pair - 1) & 1], ((pair - 1) >> 1) & 1);
              const int piece = qc == 0 ? 0 : 1;
              load_piece(next_item, piece, &tma_bar[piece], smem);
            }
            if ((qc & 1) || last_qc) {        // a pair of chunks (<= 128 queries) is complete
              const int qt = qc >> 1;
              const int wp16 = last_qc ? wp16_last : 8;
              const uint32_t tiles = sRing + (pair & 1) * 2 * TILE;
              {                               // dK[kt] += dS^T Q[pair]
                const uint64_t a = desc_at(dk0, tiles), b = mQ + qt * 1024;
#pragma unroll
                for (int j = 0; j < 8; ++j)
                  if (leader && j < wp16) umma_bf16(tDK, a + (j >> 2) * (TILE >> 4) + (j & 3) * 2, b + 128 * j, idesc_dv, (qt | j) != 0);
              }
              if (kt == 0 && qt == 0 && k > 0) mbar_wait(dq_free, (k - 1) & 1);   // dQ of the previous item read
              {                               // dQ[pair] += dS K[kt]   (A MN-major: 2 blocks of 64 queries)
                const uint64_t a = desc_at(dq0, tiles), b = mK + kt * 1024;
#pragma unroll
                for (int j = 0; j < 8; ++j)
                  if (leader && j < ksteps) umma_bf16(tDQ + 64 * qt, a + 128 * j, b + 128 * j, idesc_dq, (kt | j) != 0);
              }
              if (leader) umma_commit(&pair_done[pair & 1]);
              ++pair;
            }
            if (last_qc) ++tile_ctr;
            if (leader) BPROF(1, cc, 3);
          }
        }
        if (k + 1 < n_my) {
          if (nkt == 2) start_item(k + 1);    // queued behind this item's MMAs: no bubble at the item boundary
          mbar_wait(&pair_done[(pair - 1) & 1], ((pair - 1) >> 1) & 1);           // every operand of item k is dead
          if (nkt == 2) {
            load_piece(next_item, 2, &tma_bar[2], smem);
          } else if (k + 2 < n_my) {
            load_piece(first + (k + 2) * stride, 0, &tma_bar[k & 1], smem + (k & 1) * 4 * op_bytes);
          }
        }
      }
    }
  } else if (warp > kSoftmaxWarps) {
    // ================= vector warps: lse2[q] = lse[q]*log2(e) and D[q] = <dO[q], O[q]> of every item =================
    // 64 KB of O / dO rows per item through plain loads: done by warps of their own, an item ahead of the elementwise
    // warps (double-buffered vectors), so that nobody on the MMA <-> elementwise critical path waits for global memory.
    // A token's 128-byte head row is read by 8 lanes (16 bytes each): a warp instruction covers 4 whole rows.
    const int vw = warp - kSoftmaxWarps - 1, sub = lane & 7;
    const uint32_t vec = smem_u32(vec_base);
    constexpr int PER = 256 / kVecWarps;      // tokens per vector warp
    for (int k = 1; k < n_my; ++k) {          // (item 0: by the elementwise warps, which have nothing else to do yet)
      if (k >= 2) mbar_wait(&vec_free[k & 1], ((k >> 1) - 1) & 1);
      int cg, b, h;
      item_bh(first + k * stride, cg, b, h);
      const __nv_bfloat16* po = G.o[cg] + static_cast<size_t>(b) * N * d + h * HD + sub * 8;
      const __nv_bfloat16* pg = G.d_o[cg] + static_cast<size_t>(b) * N * d + h * HD + sub * 8;
      const float* pl = G.lse[cg] + (static_cast<size_t>(b) * H + h) * N;
      const uint32_t vl = vec + (k & 1) * 2048, vd = vl + 1024;
#pragma unroll 1
      for (int t0 = vw * PER; t0 < (vw + 1) * PER; t0 += 32) {     // 32 tokens: 8 loads per array and lane in flight
        uint4 a[8], g[8];
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          const int tok = t0 + 4 * i + (lane >> 3);
          a[i] = make_uint4(0u, 0u, 0u, 0u);
          g[i] = a[i];
          if (tok < N) {
            a[i] = *reinterpret_cast<const uint4*>(po + static_cast<size_t>(tok) * d);
            g[i] = *reinterpret_cast<const uint4*>(pg + static_cast<size_t>(tok) * d);
          }
        }
        const float l2 = t0 + lane < N ? pl[t0 + lane] * 1.4426950408889634f : INFINITY;   // padding: P = exp2(-inf) = 0
        float dd[8];
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          const uint32_t aw[4] = {a[i].x, a[i].y, a[i].z, a[i].w}, gw[4] = {g[i].x, g[i].y, g[i].z, g[i].w};
          float acc = 0.f;
#pragma unroll
          for (int e = 0; e < 4; ++e) {
            acc = fmaf(__uint_as_float(aw[e] << 16), __uint_as_float(gw[e] << 16), acc);
            acc = fmaf(__uint_as_float(aw[e] & 0xFFFF0000u), __uint_as_float(gw[e] & 0xFFFF0000u), acc);
          }
          dd[i] = acc;
        }
#pragma unroll
        for (int m = 1; m < 8; m <<= 1) {
#pragma unroll
          for (int i = 0; i < 8; ++i) dd[i] += __shfl_xor_sync(0xffffffffu, dd[i], m);
        }
        if (sub == 0) {
#pragma unroll
          for (int i = 0; i < 8; ++i) st_shared_f32(vd + (t0 + 4 * i + (lane >> 3)) * 4, dd[i]);
        }
        st_shared_f32(vl + (t0 + lane) * 4, l2);
      }
      __syncwarp();
      if (lane == 0) mbar_arrive(&vec_full[k & 1]);
    }
  } else {
    // ================= elementwise warps =================
    const int quarter = warp & 3, grp = warp >> 2;
    const int row = quarter * 32 + lane;      // key row inside the tile == TMEM lane
    const uint32_t lane_off = static_cast<uint32_t>(quarter * 32) << 16;
    const uint32_t vec = smem_u32(vec_base);
    const float sc = scale * 1.4426950408889634f;
    // 64 accumulator columns of this thread's row -> 16 per warp group -> bf16 -> 32 bytes of dqkv.
    // The rounded values that were stored are returned in v (0 for rows past the sequence) for the bias gradient.
    auto store_acc = [&](__nv_bfloat16* dqkv, uint32_t tcol, float mul, int tok, int b, int col, float (&v)[16], bool accumulate_v) {
      float f[16];
      tmem_ld_32x16(tcol + lane_off + grp * 16, f);
      tmem_ld_wait();
      uint32_t pk[8];
#pragma unroll
      for (int i = 0; i < 8; ++i) pk[i] = pack2(f[2 * i] * mul, f[2 * i + 1] * mul);
      if (tok < N) {
        uint4* dst = reinterpret_cast<uint4*>(dqkv + (static_cast<size_t>(b) * N + tok) * d3 + col + grp * 16);
        dst[0] = make_uint4(pk[0], pk[1], pk[2], pk[3]);
        dst[1] = make_uint4(pk[4], pk[5], pk[6], pk[7]);
      }
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        const float lo = tok < N ? __uint_as_float(pk[i] << 16) : 0.f, hi = tok < N ? __uint_as_float(pk[i] & 0xFFFF0000u) : 0.f;
        v[2 * i] = accumulate_v ? v[2 * i] + lo : lo;
        v[2 * i + 1] = accumulate_v ? v[2 * i + 1] + hi : hi;
      }
    };
    // staged form: the 32 bytes go into the swizzled staging tile `tile` (row = this thread's accumulator row); one
    // TMA store per tile then writes 128 rows x 128 bytes.  (Per-thread 32-byte global stores cost the LSU one
    // request per sector: ~4-6 k cycles per key tile with every elementwise warp — and the tensor pipe — waiting.)
    auto stage_acc = [&](uint32_t tile, uint32_t tcol, float mul, int tok, float (&v)[16], bool accumulate_v) {
      float f[16];
      tmem_ld_32x16(tcol + lane_off + grp * 16, f);
      tmem_ld_wait();
      uint32_t pk[8];
#pragma unroll
      for (int i = 0; i < 8; ++i) pk[i] = pack2(f[2 * i] * mul, f[2 * i + 1] * mul);
      const uint32_t rsw = (tile + row * 128) | ((row & 7) << 4);
      sts_u4(rsw ^ ((2 * grp) << 4), pk[0], pk[1], pk[2], pk[3]);
      sts_u4(rsw ^ ((2 * grp + 1) << 4), pk[4], pk[5], pk[6], pk[7]);
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        const float lo = tok < N ? __uint_as_float(pk[i] << 16) : 0.f, hi = tok < N ? __uint_as_float(pk[i] & 0xFFFF0000u) : 0.f;
        v[2 * i] = accumulate_v ? v[2 * i] + lo : lo;
        v[2 * i + 1] = accumulate_v ? v[2 * i + 1] + hi : hi;
      }
    };
    auto stage_sync = [&]() { asm volatile("bar.sync 2, 512;" ::: "memory"); };
    // colsum[col + 16*grp + c] += sum over the warp's 32 rows of v[c]: halving butterfly (16 shuffles for the 16
    // columns) and one atomic per column and warp.
    auto add_colsum = [&](float (&v)[16], float* colsum, int col) {
#pragma unroll
      for (int w = 8; w >= 1; w >>= 1) {      // lane-mask 16, 8, 4, 2: keep one half of the columns, add the partner's
        const bool hi = (lane & (2 * w)) != 0;
#pragma unroll
        for (int j = 0; j < w; ++j) {
          const float send = hi ? v[j] : v[j + w];
          const float keep = hi ? v[j + w] : v[j];
          v[j] = keep + __shfl_xor_sync(0xffffffffu, send, 2 * w);
        }
      }
      v[0] += __shfl_xor_sync(0xffffffffu, v[0], 1);
      if ((lane & 1) == 0) atomicAdd(colsum + col + grp * 16 + (lane >> 1), v[0]);
    };
    // V bias gradient without a reduction: sum_k dV[k,:] = sum_q (sum_k P[q,k]) dO[q,:] = sum_q dO[q,:].  When the
    // last key tile has a spare row (N % 128 != 0) its row 127 of P^T is set to 1 for every real query, so that
    // row of the dV accumulator IS the column sum (fp32, exact) and its 4 threads add it to dbias.
    const bool ones_row = has_dbias && (N & 127) != 0;
    if (n_my > 0) {
      // lse2 / D of the CTA's first item, by all 512 threads in one round trip (the vector warps take items 1, 2, ...)
      const int t = threadIdx.x, sub = t & 7;
      int cg, b, h;
      item_bh(first, cg, b, h);
      const __nv_bfloat16* po = G.o[cg] + static_cast<size_t>(b) * N * d + h * HD + sub * 8;
      const __nv_bfloat16* pg = G.d_o[cg] + static_cast<size_t>(b) * N * d + h * HD + sub * 8;
      uint4 a[4], g[4];
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const int tok = i * 64 + (t >> 3);
        a[i] = make_uint4(0u, 0u, 0u, 0u);
        g[i] = a[i];
        if (tok < N) {
          a[i] = *reinterpret_cast<const uint4*>(po + static_cast<size_t>(tok) * d);
          g[i] = *reinterpret_cast<const uint4*>(pg + static_cast<size_t>(tok) * d);
        }
      }
      float l2 = INFINITY;
      if (t < N) l2 = G.lse[cg][(static_cast<size_t>(b) * H + h) * N + t] * 1.4426950408889634f;
      float dd[4];
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const uint32_t aw[4] = {a[i].x, a[i].y, a[i].z, a[i].w}, gw[4] = {g[i].x, g[i].y, g[i].z, g[i].w};
        float acc = 0.f;
#pragma unroll
        for (int e = 0; e < 4; ++e) {
          acc = fmaf(__uint_as_float(aw[e] << 16), __uint_as_float(gw[e] << 16), acc);
          acc = fmaf(__uint_as_float(aw[e] & 0xFFFF0000u), __uint_as_float(gw[e] & 0xFFFF0000u), acc);
        }
        dd[i] = acc;
      }
#pragma unroll
      for (int m = 1; m < 8; m <<= 1) {
#pragma unroll
        for (int i = 0; i < 4; ++i) dd[i] += __shfl_xor_sync(0xffffffffu, dd[i], m);
      }
      if (sub == 0) {
#pragma unroll
        for (int i = 0; i < 4; ++i) st_shared_f32(vec + 1024 + (i * 64 + (t >> 3)) * 4, dd[i]);
      }
      if (t < 256) st_shared_f32(vec + t * 4, l2);
    }
    stage_sync();                             // item 0's vectors are visible to every elementwise warp
    int cc = 0, pair = 0;
    int pend_kv = -1, pend_kv_pair = 0, pend_kv_item = 0;   // key tile whose dV/dK still sit in TMEM (-1: none)
    int pend_dq_item = -1, pend_dq_pair = 0;                // item whose dQ still sits in TMEM
    auto flush_pending = [&]() {
      if (pend_kv >= 0) {
        mbar_wait(&pair_done[pend_kv_pair & 1], (pend_kv_pair >> 1) & 1);
        tc_fence_after();
        int cg, b, h;
        item_bh(pend_kv_item, cg, b, h);
        float* dbias = G.dbias[cg];
        __nv_bfloat16* dqkv = G.dqkv[cg];
        const int key = pend_kv * 128 + row;
        float v[16];
        if (ones_row && pend_kv == nkt - 1 && quarter == 3) {    // warp-uniform: tcgen05.ld is a whole-warp instruction
          float f[16];
          tmem_ld_32x16(tDV + lane_off + grp * 16, f);
          tmem_ld_wait();
          if (lane == 31) {                   // row 127, the all-ones row: sum_q dO[q, 16*grp .. +16)
#pragma unroll
            for (int c = 0; c < 16; ++c) atomicAdd(dbias + 2 * d + h * HD + grp * 16 + c, f[c]);
          }
        }
        if (staged) {
          if (threadIdx.x == 0) BPROF(0, cc, 8);
          if (threadIdx.x == 0) tma_store_wait_read();           // the staging tiles' previous stores have been read out
          if (threadIdx.x == 0) BPROF(0, cc, 9);
          stage_sync();
          if (threadIdx.x == 0) BPROF(0, cc, 10);
          stage_acc(sStage, tDV, 1.0f, key, v, false);
          if (dbias != nullptr && !ones_row) add_colsum(v, dbias, 2 * d + h * HD);
          stage_acc(sStage + TILE, tDK, scale, key, v, false);
          tc_fence_before();
          __syncwarp();
          if (lane == 0) mbar_arrive(acc_free);
          // an item that is a single chunk (N <= 64) never touches ring tile 1: dQ is staged there and leaves with dV/dK
          constexpr bool one_shot = ONE_CHUNK;
          if constexpr (one_shot) {
            stage_acc(sRing + TILE, tDQ, scale, row, v, false);
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(dq_free);
            if (dbias != nullptr) add_colsum(v, dbias, h * HD);
            pend_dq_item = -1;
          }
          if (threadIdx.x == 0) BPROF(0, cc, 11);
          fence_proxy_async();
          stage_sync();
          if (threadIdx.x == 0) BPROF(0, cc, 12);
          if (threadIdx.x == 0) {
            tma_store_3d(&G.maps[cg].dqkv_st, stage_ptr, 2 * d + h * HD, pend_kv * 128, b);
            tma_store_3d(&G.maps[cg].dqkv_st, stage_ptr + TILE, d + h * HD, pend_kv * 128, b);
            if (one_shot) tma_store_3d(&G.maps[cg].dqkv_st, stage_ptr - (kRing - 1) * TILE, h * HD, 0, b);
            tma_store_commit();
          }
          pend_kv = -1;
          return;                             // (several chunks: dQ leaves one chunk later, when these stores have drained)
        } else {
          store_acc(dqkv, tDV, 1.0f, key, b, 2 * d + h * HD, v, false);
          if (dbias != nullptr && !ones_row) add_colsum(v, dbias, 2 * d + h * HD);
          // the K bias gradient is identically zero (softmax is invariant to a shift of the scores): nothing to add
          store_acc(dqkv, tDK, scale, key, b, d + h * HD, v, false);
          tc_fence_before();
          __syncwarp();
          if (lane == 0) mbar_arrive(acc_free);
          pend_kv = -1;
        }
      }
      if (pend_dq_item >= 0) {
        mbar_wait(&pair_done[pend_dq_pair & 1], (pend_dq_pair >> 1) & 1);
        tc_fence_after();
        int cg, b, h;
        item_bh(pend_dq_item, cg, b, h);
        float* dbias = G.dbias[cg];
        float v[16];                          // both query tiles summed per thread: one reduction per item
        if (staged) {
          if (threadIdx.x == 0) tma_store_wait_read();
          stage_sync();
          for (int qt = 0; qt < nkt; ++qt) stage_acc(sStage + qt * TILE, tDQ + 64 * qt, scale, qt * 128 + row, v, qt > 0);
        } else {
          for (int qt = 0; qt < nkt; ++qt) store_acc(G.dqkv[cg], tDQ + 64 * qt, scale, qt * 128 + row, b, h * HD, v, qt > 0);
        }
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(dq_free);
        if (dbias != nullptr) add_colsum(v, dbias, h * HD);
        if (staged) {
          fence_proxy_async();
          stage_sync();
          if (threadIdx.x == 0) {
            for (int qt = 0; qt < nkt; ++qt) tma_store_3d(&G.maps[cg].dqkv_st, stage_ptr + qt * TILE, h * HD, qt * 128, b);
            tma_store_commit();
          }
        }
        pend_dq_item = -1;
      }
    };
    for (int k = 0; k < n_my; ++k) {
      const int item = first + k * stride;
      const uint32_t vl = vec + (k & 1) * 2048, vd = vl + 1024;
      // this item's lse2 / D (written by the vector warps, an item ahead); buffer 0's first fill is item 0, above
      if (k > 0) mbar_wait(&vec_full[k & 1], (k & 1) ? (k >> 1) & 1 : ((k >> 1) - 1) & 1);
      for (int kt = 0; kt < nkt; ++kt) {
        for (int qc = 0; qc < nqc; ++qc, ++cc) {
          const int W = min(64, NK - 64 * qc);
          const uint32_t tb = tmem + (cc & 1) * 128 + lane_off;
          // ring tile of this chunk; its previous user (pair - 2) must have been consumed by dK/dQ
          if (threadIdx.x == 0) BPROF(0, cc, 0);
          if ((qc & 1) == 0 && pair >= 2) mbar_wait(&pair_done[pair & 1], ((pair >> 1) - 1) & 1);
          if (threadIdx.x == 0) BPROF(0, cc, 1);
          mbar_wait(&sdp_full[cc & 1], (cc >> 1) & 1);
          tc_fence_after();
          if (threadIdx.x == 0) BPROF(0, cc, 2);
          if (16 * grp < W) {                 // warp-uniform
            float s[16], dp[16];
            tmem_ld_32x16(tb + 16 * grp, s);
            tmem_ld_32x16(tb + 64 + 16 * grp, dp);
            tmem_ld_wait();
            if (threadIdx.x == 0) BPROF(0, cc, 3);
            const int q0 = 64 * qc + 16 * grp;
            uint32_t pp[8], ds[8];
#pragma unroll
            for (int i = 0; i < 16; i += 4) {
              float l[4], dv[4];
              ld_shared_f32x4(vl + (q0 + i) * 4, l);
              ld_shared_f32x4(vd + (q0 + i) * 4, dv);
              float p[4], g[4];
#pragma unroll
              for (int e = 0; e < 4; ++e) {
                p[e] = ex2(fmaf(s[i + e], sc, -l[e]));
                g[e] = p[e] * (dp[i + e] - dv[e]);
              }
              if (ones_row && row == 127 && kt == nkt - 1) {     // see ones_row: P^T[127, q] = [q < N]
#pragma unroll
                for (int e = 0; e < 4; ++e) p[e] = q0 + i + e < N ? 1.0f : 0.0f;
              }
              pp[i >> 1] = pack2(p[0], p[1]);
              pp[(i >> 1) + 1] = pack2(p[2], p[3]);
              ds[i >> 1] = pack2(g[0], g[1]);
              ds[(i >> 1) + 1] = pack2(g[2], g[3]);
            }
            tmem_st_32x8(tb + 16 * grp, pp);  // P^T over this thread's own (consumed) scores
            const uint32_t tile = sRing + ((pair & 1) * 2 + (qc & 1)) * TILE;
            const uint32_t rsw = (tile + row * 128) | ((row & 7) << 4);
            sts_u4(rsw ^ ((2 * grp) << 4), ds[0], ds[1], ds[2], ds[3]);
            sts_u4(rsw ^ ((2 * grp + 1) << 4), ds[4], ds[5], ds[6], ds[7]);
            tmem_st_wait();
            fence_proxy_async();
          }
          if (threadIdx.x == 0) BPROF(0, cc, 4);
          tc_fence_before();
          __syncwarp();
          if (lane == 0) mbar_arrive(&e_done[cc & 1]);
          if (threadIdx.x == 0) BPROF(0, cc, 5);
          flush_pending();                    // outputs of the previous key tile / item, now that this chunk is handed over
          if (threadIdx.x == 0) BPROF(0, cc, 6);
          if ((qc & 1) || qc == nqc - 1) {
            if (qc == nqc - 1) {
              pend_kv = kt;
              pend_kv_pair = pair;
              pend_kv_item = item;
              if (kt == nkt - 1) {
                pend_dq_item = item;
                pend_dq_pair = pair;
              }
            }
            ++pair;
          }
          if (threadIdx.x == 0) BPROF(0, cc, 7);
        }
      }
      __syncwarp();
      if (lane == 0) mbar_arrive(&vec_free[k & 1]);
    }
    flush_pending();
    flush_pending();                          // (staged: dQ leaves in a second step)
    if (staged && threadIdx.x == 0) tma_store_wait_all();
  }
  tc_fence_before();
  __syncthreads();
  if (warp == kSoftmaxWarps) {
    tc_fence_after();
    tmem_dealloc(tmem, 512);
  }
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
EncodeTiledFn encode_fn() {
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
// bf16 [B, N, width] row-major; box = 64 columns x `rows` token rows x 1 sample, 128-byte swizzle
int make_tmap3(CUtensorMap* m, const void* ptr, int B, int N, int width, int rows) {
  EncodeTiledFn fn = encode_fn();
  if (!fn) FC_FAIL(FC_ERR_CUDA, "cuTensorMapEncodeTiled entry point unavailable");
  cuuint64_t gdim[3] = {(cuuint64_t)width, (cuuint64_t)N, (cuuint64_t)B};
  cuuint64_t gstr[2] = {(cuuint64_t)width * 2, (cuuint64_t)N * width * 2};
  cuuint32_t box[3] = {64, (cuuint32_t)rows, 1};
  cuuint32_t estr[3] = {1, 1, 1};
  CUresult r = fn(m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 3, const_cast<void*>(ptr), gdim, gstr, box, estr,
                  CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                  CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) FC_FAIL(FC_ERR_CUDA, "cuTensorMapEncodeTiled (3d) failed (%d)", (int)r);
  return FC_OK;
}

}  // namespace

extern "C" int fc_attention_fwd_grouped(int groups, const void* const* qkv, void* const* out, float* const* lse, int B,
                                        int N, int H, int head_dim, int device, void* stream) {
  FC_REQUIRE(groups >= 1 && groups <= MAXG, "fc_attention_fwd: %d groups (1..%d)", groups, MAXG);
  FC_REQUIRE(head_dim == HD, "fc_attention_fwd: head_dim must be 64 (got %d)", head_dim);
  FC_REQUIRE(B > 0 && N > 0 && N <= 256 && H > 0 && static_cast<long long>(B) * H < 65536,
             "fc_attention_fwd: unsupported shape B=%d N=%d H=%d", B, N, H);
  FcDeviceGuard guard(device);
  const int RA = N > 128 ? 128 : ((N + 15) & ~15);
  const int RB = N > 128 ? ((N - 128 + 15) & ~15) : 0;
  const int NK = RA + RB;
  FwdGroups G;
  memset(&G, 0, sizeof(G));
  for (int g = 0; g < groups; ++g) {
    FC_REQUIRE(qkv[g] != nullptr && out[g] != nullptr, "fc_attention_fwd: null tensor (group %d)", g);
    FC_REQUIRE((reinterpret_cast<uintptr_t>(qkv[g]) & 15) == 0 && (reinterpret_cast<uintptr_t>(out[g]) & 15) == 0,
               "fc_attention_fwd: qkv/out must be 16-byte aligned");
    int rc = make_tmap3(&G.maps[g].qkv_a, qkv[g], B, N, 3 * H * HD, RA);
    if (!rc) rc = make_tmap3(&G.maps[g].qkv_b, qkv[g], B, N, 3 * H * HD, RB ? RB : RA);
    if (!rc) rc = make_tmap3(&G.maps[g].out_st, out[g], B, N, H * HD, RA);
    if (rc) return rc;
    G.out[g] = static_cast<__nv_bfloat16*>(out[g]);
    G.lse[g] = lse ? lse[g] : nullptr;
  }
  const int buf_bytes = 3 * NK * 128;
  // O leaves through two staging tiles + TMA stores when two operand buffers still fit beside them
  const int staged = 2 * buf_bytes + 6144 + 128 + 2 * TILE <= kMaxDynSmem;
  const int aux = 6144 + 128 + (staged ? 2 * TILE : 0);
  int nbuf = 4;                                                  // operand buffers: as many as fit (<= 4)
  while (nbuf > 1 && nbuf * buf_bytes + aux > kMaxDynSmem) --nbuf;
  // QK^T is issued with M = 128 whatever N is: the tensor core reads 128 rows (16 KB) from the Q tile's address.  Rows
  // past the item's NK only produce score rows nobody reads, but for very short sequences (NK = 16: a 6 KB buffer)
  // the read from the LAST buffer would run past the CTA's shared-memory allocation -> pad the allocation.
  const int pad = buf_bytes < TILE ? TILE - buf_bytes : 0;
  const int smem = nbuf * buf_bytes + aux + pad;
  FC_SMEM_OPT_IN(attn_fwd_tc_kernel<true>, kMaxDynSmem);
  FC_SMEM_OPT_IN(attn_fwd_tc_kernel<false>, kMaxDynSmem);
  const int items = groups * B * H;
  const int sms = fc_num_sms(device);
  const int waves = (items + sms - 1) / sms;
  const int grid = fc_apply_grid_cap((items + waves - 1) / waves);   // balanced persistent grid (<= #SMs)
  const float scale_log2e = 0.125f * 1.4426950408889634f;        // 64^-0.5 * log2(e)
  if (staged)
    attn_fwd_tc_kernel<true><<<grid, kFwdThreads, smem, reinterpret_cast<cudaStream_t>(stream)>>>(G, items, B * H, N, H, nbuf, scale_log2e);
  else
    attn_fwd_tc_kernel<false><<<grid, kFwdThreads, smem, reinterpret_cast<cudaStream_t>(stream)>>>(G, items, B * H, N, H, nbuf, scale_log2e);
  FC_LAUNCH_CHECK();
  return FC_OK;
}

extern "C" int fc_attention_fwd(const void* qkv, void* out, float* lse, int B, int N, int H, int head_dim,
                                int device, void* stream) {
  return fc_attention_fwd_grouped(1, &qkv, &out, &lse, B, N, H, head_dim, device, stream);
}

extern "C" int fc_attention_bwd_grouped(int groups, const void* const* qkv, const void* const* out,
                                        const void* const* d_out, const float* const* lse, void* const* dqkv,
                                        float* const* dbias, int B, int N, int H, int head_dim, int device, void* stream) {
  FC_REQUIRE(groups >= 1 && groups <= MAXG, "fc_attention_bwd: %d groups (1..%d)", groups, MAXG);
  FC_REQUIRE(head_dim == HD, "fc_attention_bwd: head_dim must be 64 (got %d)", head_dim);
  FC_REQUIRE(B > 0 && N > 0 && N <= 256 && H > 0 && static_cast<long long>(B) * H < 65536,
             "fc_attention_bwd: unsupported shape B=%d N=%d H=%d", B, N, H);
  FcDeviceGuard guard(device);
  const int RA = N > 128 ? 128 : ((N + 15) & ~15);
  const int RB = N > 128 ? ((N - 128 + 15) & ~15) : 0;
  const int NK = RA + RB;
  BwdGroups G;
  memset(&G, 0, sizeof(G));
  for (int g = 0; g < groups; ++g) {
    FC_REQUIRE(qkv[g] && out[g] && d_out[g] && lse[g] && dqkv[g], "fc_attention_bwd: null tensor (group %d)", g);
    FC_REQUIRE(((reinterpret_cast<uintptr_t>(qkv[g]) | reinterpret_cast<uintptr_t>(out[g]) |
                 reinterpret_cast<uintptr_t>(d_out[g]) | reinterpret_cast<uintptr_t>(dqkv[g])) & 15) == 0,
               "fc_attention_bwd: tensors must be 16-byte aligned");
    FC_REQUIRE(((dbias ? dbias[g] : nullptr) == nullptr) == ((dbias ? dbias[0] : nullptr) == nullptr),
               "fc_attention_bwd: groups disagree on dbias");
    BwdMaps& m = G.maps[g];
    int rc = make_tmap3(&m.qkv_a, qkv[g], B, N, 3 * H * HD, RA);
    if (!rc) rc = make_tmap3(&m.qkv_b, qkv[g], B, N, 3 * H * HD, RB ? RB : RA);
    if (!rc) rc = make_tmap3(&m.do_a, d_out[g], B, N, H * HD, RA);
    if (!rc) rc = make_tmap3(&m.do_b, d_out[g], B, N, H * HD, RB ? RB : RA);
    if (!rc) rc = make_tmap3(&m.dqkv_st, dqkv[g], B, N, 3 * H * HD, RA);
    if (rc) return rc;
    G.o[g] = static_cast<const __nv_bfloat16*>(out[g]);
    G.d_o[g] = static_cast<const __nv_bfloat16*>(d_out[g]);
    G.lse[g] = lse[g];
    G.dqkv[g] = static_cast<__nv_bfloat16*>(dqkv[g]);
    G.dbias[g] = dbias ? dbias[g] : nullptr;
  }
  int smem = (RB ? 1 : 2) * 4 * NK * 128 + kRing * TILE + 4096 + 256;            // two operand sets for one key tile
  // gradients leave through two staging tiles + TMA stores when they fit
  const int staged = smem + 2 * TILE <= kMaxDynSmem ? 2 : 0;
  smem += staged * TILE;
  FC_SMEM_OPT_IN(attn_bwd_tc_kernel<false>, kMaxDynSmem);   // one process-wide value: the attribute is per function, not per thread
  FC_SMEM_OPT_IN(attn_bwd_tc_kernel<true>, kMaxDynSmem);
  const int items = groups * B * H;
  const int sms = fc_num_sms(device);
  const int waves = (items + sms - 1) / sms;
  const int grid = fc_apply_grid_cap((items + waves - 1) / waves);
  if (NK <= 64 && staged)
    attn_bwd_tc_kernel<true><<<grid, kBwdThreads, smem, reinterpret_cast<cudaStream_t>(stream)>>>(G, items, B * H, N, H, 0.125f, staged);
  else
    attn_bwd_tc_kernel<false><<<grid, kBwdThreads, smem, reinterpret_cast<cudaStream_t>(stream)>>>(G, items, B * H, N, H, 0.125f, staged);
  FC_LAUNCH_CHECK();
  return FC_OK;
}

extern "C" int fc_attention_bwd(const void* qkv, const void* out, const void* d_out, const float* lse, void* dqkv,
                                float* dbias, int B, int N, int H, int head_dim, int device, void* stream) {
  void* dq = dqkv;
  return fc_attention_bwd_grouped(1, &qkv, &out, &d_out, &lse, &dq, &dbias, B, N, H, head_dim, device, stream);
}
