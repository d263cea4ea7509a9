echo "=== gemm tests"; timeout 300 python -m pytest tests/test_gemm_gpu.py tests/test_multitile_gpu.py -m gpu -q -p no:cacheprovider -k "gemm or patch" 2>&1 | tail -4
timeout 200 python tools/gemm_group_bench.py 2>&1 | tail -14
timeout 200 python tools/gemm_group_bench.py 7168 2>&1 | tail -14
FC_GEMM_PAIR=1 timeout 200 python tools/gemm_group_bench.py 2>&1 | tail -14
