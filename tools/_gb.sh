for m in 0 1 2; do echo "=== FC_GEMM_DEBUG=$m"; FC_GEMM_DEBUG=$m timeout 120 python tools/gemm_bench.py 2>&1 | tail -12; done
