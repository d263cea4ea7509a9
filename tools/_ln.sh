echo "=== tests"; timeout 300 python -m pytest tests/test_kernels_gpu.py -m gpu -q -p no:cacheprovider -k layernorm 2>&1 | tail -4
for r in 0 1; do echo "=== ln_bench FC_LN_RING=$r"; FC_LN_RING=$r timeout 100 python tools/ln_bench.py 2>&1 | tail -3; done
