mkdir -p gpurun_out
timeout 600 python tools/bench_gemm.py linear_sweep 2>&1 | tee gpurun_out/linear_sweep.txt | tail -120
