for cg in 1 3; do
  timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --cache-control none --profile-from-start off --csv --log-file gpurun_out/r2_launches_cg$cg.csv python bench.py --profile --threads 1 --client-group $cg > /dev/null 2> gpurun_out/ncu_cg$cg.err
  python tools/summarize_launches.py gpurun_out/r2_launches_cg$cg.csv > gpurun_out/r2_launches_cg${cg}_summary.txt 2>&1
  head -30 gpurun_out/r2_launches_cg${cg}_summary.txt
done
