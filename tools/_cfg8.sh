N=${1:-8}
run() { name=$1; shift; echo "=== $name N=$N"; timeout 240 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus $N --steps 3 --warmup 2 --no-e2e "$@" 2>gpurun_out/cfg8_$name.err | tee gpurun_out/r2_cfg_${name}_n$N.json | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['value'], d['ms_per_step'], d['scaling'], d['roofline']['frac'], d['roofline']['round_model_tflops'], d['aggregation']['ms'], d['per_round_ms'])"; grep -v "^  File\|^    \|^\*\*\*\|OMP_NUM" gpurun_out/cfg8_$name.err | tail -2; }
run vits-coco-reference --config vits-coco --placement reference
run vits-coco-balanced --config vits-coco --placement balanced
run vitb-fediot --config vitb-fediot --placement balanced
run vitb-fedprox --config vitb-fedprox --placement balanced
