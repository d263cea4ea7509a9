// Stand-alone phase profile of the tcgen05 attention backward (clock64 stamps of CTA 0 / thread 0, per chunk).
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -DFC_ATTN_PROF -Ifedcola_b200/csrc \
//        tools/attn_prof_bwd.cu fedcola_b200/csrc/api.cu -lcuda -o tools/attn_prof_bwd.bin
#include "../fedcola_b200/csrc/attention.cu"
#include <cstdio>
#include <vector>
extern "C" int fc_colsum_bf16(const void*, long long, int, int, float*, int, void*) { return 0; }
int main(int argc, char** argv) {
  int B = 112, N = argc > 1 ? atoi(argv[1]) : 176, H = 6;
  size_t n = (size_t)B * N * 3 * H * 64;
  std::vector<__nv_bfloat16> h(n);
  for (size_t i = 0; i < n; ++i) h[i] = __float2bfloat16((float)((i * 2654435761u) % 1000) / 500.f - 1.f);
  __nv_bfloat16 *qkv, *out, *dout, *dqkv; float *lse, *dbias;
  cudaMalloc(&qkv, n * 2); cudaMalloc(&dqkv, n * 2); cudaMalloc(&out, n * 2 / 3); cudaMalloc(&dout, n * 2 / 3);
  cudaMalloc(&lse, (size_t)B * H * N * 4); cudaMalloc(&dbias, 3 * H * 64 * 4);
  cudaMemset(dbias, 0, 3 * H * 64 * 4);
  cudaMemcpy(qkv, h.data(), n * 2, cudaMemcpyHostToDevice);
  cudaMemcpy(dout, h.data(), n * 2 / 3, cudaMemcpyHostToDevice);
  int rc = fc_attention_fwd(qkv, out, lse, B, N, H, 64, 0, nullptr);
  for (int it = 0; it < 3 && !rc; ++it) { rc = fc_attention_bwd(qkv, out, dout, lse, dqkv, dbias, B, N, H, 64, 0, nullptr); cudaDeviceSynchronize(); }
  if (rc) { printf("error %d: %s\n", rc, fc_last_error()); return 1; }
  long long p[16 * 12];
  cudaMemcpyFromSymbol(p, g_attn_prof, sizeof(p));
  const char* names[] = {"top", "ring wait", "SdP wait", "tmem ld", "math+st issue", "st wait+fence+arrive", "flush", "vectors"};
  for (int c = 0; c < 16; ++c) {
    printf("chunk %2d:", c);
    for (int s = 1; s < 8; ++s) printf(" %s +%lld |", names[s], p[c * 12 + s] - p[c * 12 + s - 1]);
    if (c) printf("  (gap %lld)", p[c * 12] - p[(c - 1) * 12 + 7]);
    printf("  total %lld\n", p[c * 12 + 7] - p[c * 12]);
  }
  return 0;
}
