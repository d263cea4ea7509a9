timeout 600 python -m pytest tests/test_eval_gpu.py -m gpu -q -p no:cacheprovider 2>&1 | tail -30
