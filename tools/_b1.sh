timeout 900 python bench.py --gpus 1 --steps 4 --warmup 3 2> gpurun_out/b1.err | tee gpurun_out/r2_bench_1gpu_a.json | cut -c1-300; grep -v "^  File\|^    " gpurun_out/b1.err | tail -25
timeout 300 python -m pytest tests/test_model_gpu.py -m gpu -q -p no:cacheprovider -k vit_sized 2>&1 | tail -5
