export FC_GEMM_PAIR=1
echo "=== gemm tests (pair)"; timeout 200 python -m pytest tests/test_gemm_gpu.py -m gpu -q -p no:cacheprovider 2>&1 | tail -4
for bn in 0 256; do for m in 0 1; do echo "=== pair BN=$bn FC_GEMM_DEBUG=$m"; FC_GEMM_BN=$bn FC_GEMM_DEBUG=$m timeout 120 python tools/gemm_bench.py 2>&1 | tail -10; done; done
