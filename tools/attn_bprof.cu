// Stand-alone phase profile of the tcgen05 attention BACKWARD (clock64 stamps of CTA 0: elementwise warp 0 and the
// control lane, first 32 chunks).
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -DFC_ATTN_PROF -Ifedcola_b200/csrc -Iinclude \
//        tools/attn_bprof.cu fedcola_b200/csrc/api.cu -lcuda -o attn_bprof          (run: ./attn_bprof [N])
#include "../fedcola_b200/csrc/attention.cu"
#include <cstdio>
#include <vector>
int main(int argc, char** argv) {
  int B = 112, N = argc > 1 ? atoi(argv[1]) : 197, H = 6;
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
  if (rc) { printf("error %d: %s\n", rc, fc_last_error()); return 1; }
  for (int it = 0; it < 3; ++it) {
    rc = fc_attention_bwd(qkv, out, dout, lse, dqkv, dbias, B, N, H, 64, 0, nullptr);
    if (rc) { printf("error %d: %s\n", rc, fc_last_error()); return 1; }
    cudaDeviceSynchronize();
  }
  long long p[2][32 * 16];
  cudaMemcpyFromSymbol(p, g_attn_bprof, sizeof(p));
  const long long t0 = p[0][0];
  printf("elementwise warp 0 (cycles): start | wait ring | wait S,dP | tmem ld | math+st | arrive | flush | vectors || period\n");
  for (int c = 0; c < 32 && p[0][c * 16] > 0; ++c) {
    const long long* q = p[0] + c * 16;
    printf("chunk %2d @%8lld: %6lld %6lld %6lld %6lld %6lld %6lld %6lld || %6lld\n", c, q[0] - t0, q[1] - q[0], q[2] - q[1],
           q[3] ? q[3] - q[2] : 0, q[3] ? q[4] - q[3] : q[4] - q[2], q[5] - q[4], q[6] - q[5], q[7] - q[6], c ? q[0] - p[0][(c - 1) * 16] : 0);
    if (q[8] > 0) printf("          kv flush: wait pair %lld | wait_read %lld | sync %lld | stage+colsum %lld | fence+sync %lld | issue %lld\n",
                         q[8] - q[5], q[9] - q[8], q[10] - q[9], q[11] - q[10], q[12] - q[11], q[6] - q[12]);
  }
  printf("control lane (cycles): start | wait e_done | wait acc_free | issue\n");
  for (int c = 0; c < 32 && p[1][c * 16] > 0; ++c) {
    const long long* q = p[1] + c * 16;
    printf("chunk %2d @%8lld: %6lld %6lld %6lld\n", c, q[0] - t0, q[1] - q[0], q[2] - q[1], q[3] - q[2]);
  }
  return 0;
}
