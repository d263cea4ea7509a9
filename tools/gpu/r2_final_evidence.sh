# Final round-2 evidence (one B200): in-situ launch list of one bench round, ncu metrics of 240 consecutive grouped launches,
# CUPTI timeline, attention / GEMM / LayerNorm micro-benchmarks.  Outputs under gpurun_out/final/.
mkdir -p gpurun_out/final
timeout 300 python bench.py --timeline 2> /dev/null; cp gpurun_out/timeline.txt gpurun_out/final/round_timeline_cupti.txt
timeout 200 python tools/attn_bench.py > gpurun_out/final/attn_bench.log 2>&1
timeout 200 python tools/gemm_scaling.py > gpurun_out/final/gemm_fixed_vs_steady.log 2>&1
FC_GEMM_DEBUG=1 timeout 200 python tools/gemm_scaling.py >> gpurun_out/final/gemm_fixed_vs_steady.log 2>&1
timeout 200 python tools/ln_bench.py > gpurun_out/final/ln_bench.log 2>&1
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --cache-control none --profile-from-start off --csv --log-file gpurun_out/final/launches.csv python bench.py --profile --threads 1 > /dev/null 2> gpurun_out/final/ncu_l.err
python tools/summarize_launches.py gpurun_out/final/launches.csv gpurun_out/final/launches_insitu_summary.txt | head -12
timeout 600 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed,gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed,sm__throughput.avg.pct_of_peak_sustained_elapsed --clock-control none --cache-control none --profile-from-start off -k regex:'gemm_bf16|attn_|ln_' -s 400 -c 240 --csv --page raw --log-file gpurun_out/final/kernels_raw.csv python bench.py --profile --threads 1 > /dev/null 2> gpurun_out/final/ncu_k.err
python tools/ncu_kernel_stats.py gpurun_out/final/kernels_raw.csv gpurun_out/final/kernel_stats.json > gpurun_out/final/kernel_stats.txt; head -16 gpurun_out/final/kernel_stats.txt
rm -f gpurun_out/final/launches.csv gpurun_out/final/kernels_raw.csv
