// Instruction-throughput probes (cycles per warp-instruction per SM sub-partition).
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 tools/ubench.cu -o tools/ubench.bin
#include <cstdio>
#include <cstdint>
#include <cuda_bf16.h>
__device__ __forceinline__ uint32_t pack_hw(float a, float b) {
  uint32_t r;
  asm volatile("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(b), "f"(a));
  return r;
}
__device__ __forceinline__ uint32_t pack_int(float a, float b) {   // RNE by integer arithmetic + byte permute
  uint32_t ua = __float_as_uint(a), ub = __float_as_uint(b);
  ua += 0x7FFFu + ((ua >> 16) & 1u);
  ub += 0x7FFFu + ((ub >> 16) & 1u);
  return __byte_perm(ua, ub, 0x7632);
}
template <int MODE>
__global__ void probe(float* out, long long* cyc, float seed) {
  float x[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) x[i] = seed + threadIdx.x * 0.001f + i;
  uint32_t acc = 0;
  __syncthreads();
  long long t0 = clock64();
#pragma unroll 1
  for (int it = 0; it < 512; ++it) {
#pragma unroll
    for (int i = 0; i < 8; i += 2) {
      if (MODE == 0) acc ^= pack_hw(x[i], x[i + 1]);
      if (MODE == 1) acc ^= pack_int(x[i], x[i + 1]);
      if (MODE == 2) { float y; asm volatile("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x[i])); acc ^= __float_as_uint(y);
                       asm volatile("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x[i+1])); acc ^= __float_as_uint(y); }
      if (MODE == 3) { acc ^= __float_as_uint(fmaf(x[i], x[i + 1], seed)); acc ^= __float_as_uint(fmaf(x[i+1], x[i], seed)); }
      x[i] += 1.0f; x[i + 1] += 1.0f;
    }
  }
  long long t1 = clock64();
  __syncthreads();
  out[blockIdx.x * blockDim.x + threadIdx.x] = __uint_as_float(acc);
  if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}
int main() {
  float* out; long long* cyc;
  cudaMalloc(&out, 1 << 20); cudaMalloc(&cyc, 1024);
  const char* names[] = {"cvt.rn.bf16x2.f32 (4/iter)", "integer RNE + prmt (4 packs/iter)", "ex2.approx (8/iter)", "fma (8/iter) baseline"};
  for (int mode = 0; mode < 4; ++mode)
    for (int warps : {4, 16}) {
      long long h = 0;
      for (int rep = 0; rep < 2; ++rep) {
        if (mode == 0) probe<0><<<1, warps * 32>>>(out, cyc, 1.5f);
        if (mode == 1) probe<1><<<1, warps * 32>>>(out, cyc, 1.5f);
        if (mode == 2) probe<2><<<1, warps * 32>>>(out, cyc, 1.5f);
        if (mode == 3) probe<3><<<1, warps * 32>>>(out, cyc, 1.5f);
        cudaDeviceSynchronize();
        cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost);
      }
      printf("%-36s warps/SM=%2d: %6.2f cycles per loop iteration (per warp-slot: %.2f)\n", names[mode], warps, h / 512.0,
             h / 512.0 / (warps / 4));
    }
  return 0;
}
