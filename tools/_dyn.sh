echo "=== gemm tests (dyn sched default)"; timeout 300 python -m pytest tests/test_gemm_gpu.py tests/test_multitile_gpu.py -m gpu -q -p no:cacheprovider -k "gemm or patch or sched" 2>&1 | tail -6
echo "=== model tests"; timeout 300 python -m pytest tests/test_model_gpu.py -m gpu -q -p no:cacheprovider 2>&1 | tail -4
for sc in 0 1; do echo "=== bench FC_GEMM_SCHED=$sc"; FC_GEMM_SCHED=$sc timeout 200 python bench.py --gpus 1 --steps 4 --warmup 2 --no-cpu-baseline --no-parity-check --no-gpu-eager --no-e2e 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['value'], d['ms_per_step'], d['roofline']['frac'], d['roofline']['avg_launch_us'], d['per_round_ms'])"; done
FC_GEMM_SCHED=1 timeout 100 python tools/gemm_bench.py 2>&1 | tail -10
