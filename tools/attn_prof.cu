// Stand-alone phase profile of the tcgen05 attention forward (clock64 stamps of CTA 0 / thread 0).
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -DFC_ATTN_PROF -Ifedcola_b200/csrc \
//        tools/attn_prof.cu fedcola_b200/csrc/api.cu -lcuda -o gpurun_out/attn_prof
#include "../fedcola_b200/csrc/attention.cu"
#include <cstdio>
#include <vector>
int main(int argc, char** argv) {
  int B = 112, N = argc > 1 ? atoi(argv[1]) : 197, H = 6;
  size_t n = (size_t)B * N * 3 * H * 64;
  std::vector<__nv_bfloat16> h(n);
  for (size_t i = 0; i < n; ++i) h[i] = __float2bfloat16((float)((i * 2654435761u) % 1000) / 500.f - 1.f);
  __nv_bfloat16 *qkv, *out; float* lse;
  cudaMalloc(&qkv, n * 2); cudaMalloc(&out, n * 2 / 3); cudaMalloc(&lse, (size_t)B * H * N * 4);
  cudaMemcpy(qkv, h.data(), n * 2, cudaMemcpyHostToDevice);
  for (int it = 0; it < 3; ++it) {
    int rc = fc_attention_fwd(qkv, out, lse, B, N, H, 64, 0, nullptr);
    if (rc) { printf("error %d: %s\n", rc, fc_last_error()); return 1; }
    cudaDeviceSynchronize();
  }
#ifdef FC_ATTN_PROF
  long long p[16 * 12];
  cudaMemcpyFromSymbol(p, g_attn_prof, sizeof(p));
  const char* names[] = {"tile start", "S ready", "-", "max+sync", "sums+prev O ld", "exp+P+arrive", "O stg"};
  for (int T = 0; T < 9; ++T) {
    printf("tile %d:", T);
    for (int s = 1; s < 7; ++s) if (s != 2) printf(" %s +%lld |", names[s], p[T * 12 + s] - p[T * 12 + (s == 3 ? 1 : s - 1)]);
    if (T) printf("  (gap from prev %lld)", p[T * 12] - p[(T - 1) * 12 + 6]);
    printf("  total %lld\n", p[T * 12 + 6] - p[T * 12]);
  }
#endif
  return 0;
}
