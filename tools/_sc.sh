for pair in 0 1; do for m in 0 1; do FC_GEMM_PAIR=$pair FC_GEMM_DEBUG=$m timeout 200 python tools/gemm_scaling.py 2>&1 | tail -7; done; done
