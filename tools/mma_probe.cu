// tcgen05.mma cost per instruction as a function of the tile shape: one CTA, one issuing thread, R back-to-back
// kind::f16 MMAs (M = 128, K = 16) of width N on zeroed shared memory, SS (A from smem) and TS (A from TMEM) forms;
// optionally rotating over several accumulators (dependent vs independent MMAs);
// cycles from the first issue to the commit's arrival, and the cycles the issuing thread spent in the issue loop.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -Ifedcola_b200/csrc -Iinclude tools/mma_probe.cu -lcuda -o mma_probe
#include <cstdio>
#include <cstdint>
#include <cuda.h>
#include <cuda_bf16.h>
#include "../fedcola_b200/csrc/sm100.cuh"
using namespace sm100;

template <int ND, int UNROLL, int MODE>
__global__ void __launch_bounds__(128, 1) probe(long long* out, int N, int R, int ts, int b_mn) {
  extern __shared__ __align__(1024) uint8_t smem[];
  __shared__ uint64_t bar;
  __shared__ uint32_t slot;
  for (int i = threadIdx.x; i < 96 * 1024 / 4; i += blockDim.x) reinterpret_cast<uint32_t*>(smem)[i] = 0;
  if (threadIdx.x == 0) { mbar_init(&bar, 1); fence_barrier_init(); }
  if (threadIdx.x < 32) { tmem_alloc(&slot, 512); tmem_relinquish(); }
  fence_proxy_async();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = slot;
  if (MODE == 0 ? threadIdx.x == 0 : threadIdx.x < 32) {
    const bool leader = MODE == 0 ? true : elect_one();
    const uint32_t base = smem_u32(smem);
    const uint64_t a0 = umma_smem_desc(base, 16, 1024);                       // K-major, 128B swizzle tile
    const uint64_t b0 = b_mn ? umma_smem_desc(base + 32768, 8192, 1024) : umma_smem_desc(base + 32768, 16, 1024);
    const uint32_t idesc = umma_idesc_bf16(128, N, 0, b_mn);
    for (int rep = 0; rep < 2; ++rep) {
      const long long t0 = clock64();
      if (ts) {
#pragma unroll 1
        for (int r = 0; r < R; r += UNROLL) {
#pragma unroll
          for (int u = 0; u < UNROLL; ++u)
            if (leader) umma_bf16_ts(tmem + 256 + (u % ND) * 64, tmem + 8 * (u & 3), b0 + 2 * (u & 3), idesc, (r | (u / ND)) != 0);
        }
      } else {
#pragma unroll 1
        for (int r = 0; r < R; r += UNROLL) {
#pragma unroll
          for (int u = 0; u < UNROLL; ++u)
            if (leader) umma_bf16(tmem + 256 + (u % ND) * 64, a0 + 2 * (u & 3), b0 + 2 * (u & 3), idesc, (r | (u / ND)) != 0);
        }
      }
      const long long t1 = clock64();
      if (leader) umma_commit(&bar);
      if (MODE) __syncwarp();
      mbar_wait(&bar, rep & 1);
      const long long t2 = clock64();
      if (blockIdx.x == 0 && leader) { out[0] = t1 - t0; out[1] = t2 - t0; }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (threadIdx.x < 32) { tc_fence_after(); tmem_dealloc(tmem, 512); }
}

template <int ND, int UNROLL, int MODE = 0>
void run(long long* d, int N, int ts, int b_mn) {
  const int R = 256;
  cudaFuncSetAttribute(probe<ND, UNROLL, MODE>, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024);
  probe<ND, UNROLL, MODE><<<1, 128, 100 * 1024>>>(d, N, R, ts, b_mn);
  long long h[2];
  cudaError_t e = cudaMemcpy(h, d, 16, cudaMemcpyDeviceToHost);
  if (e != cudaSuccess) { printf("error %s\n", cudaGetErrorString(e)); exit(1); }
  printf("%s %s B %s N=%3d accumulators=%d unroll=%2d: issue loop %6.1f cyc/MMA, to completion %6.1f cyc/MMA (floor %d)\n", MODE ? "uniform+elect" : "lane0 branch  ", ts ? "TS" : "SS",
         b_mn ? "MN-major" : "K-major ", N, ND, UNROLL, (double)h[0] / R, (double)h[1] / R, 128 * N / 256);
}

int main() {
  long long* d;
  cudaMalloc(&d, 16);
  for (int ts = 0; ts < 2; ++ts)
    for (int N : {32, 64, 128, 256}) {
      run<1, 1>(d, N, ts, 0);
      run<1, 16>(d, N, ts, 0);
      run<1, 1, 1>(d, N, ts, 0);
      run<1, 4, 1>(d, N, ts, 0);
      run<1, 16, 1>(d, N, ts, 0);
      if (N <= 64) run<4, 16, 1>(d, N, ts, 0);
    }
  run<1, 16, 1>(d, 64, 0, 1);
  return 0;
}
