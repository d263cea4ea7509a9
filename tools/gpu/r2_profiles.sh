# Round-2 evidence run (one B200): aggregation sweep (BASELINE configs[4]), in-situ launch list of one bench round, and an
# ncu metrics pass over 240 consecutive GEMM / attention / LayerNorm launches of that round.
set -x
timeout 900 python bench.py --config agg-sweep 2> gpurun_out/agg_sweep.err > gpurun_out/r2_agg_sweep.json; tail -2 gpurun_out/agg_sweep.err
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --cache-control none --profile-from-start off --csv --log-file gpurun_out/r2_launches_final.csv python bench.py --profile --threads 1 > /dev/null 2> gpurun_out/ncu_l.err
python tools/summarize_launches.py gpurun_out/r2_launches_final.csv gpurun_out/r2_launches_final_summary.txt | head -16
timeout 900 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed,gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed,sm__throughput.avg.pct_of_peak_sustained_elapsed --clock-control none --cache-control none --profile-from-start off -k regex:'gemm_bf16|attn_|ln_' -s 400 -c 240 --csv --page raw --log-file gpurun_out/r2_kernels_raw.csv python bench.py --profile --threads 1 > /dev/null 2> gpurun_out/ncu_k.err
python tools/ncu_kernel_stats.py gpurun_out/r2_kernels_raw.csv gpurun_out/r2_kernel_stats.json
