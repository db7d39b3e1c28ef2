mkdir -p gpurun_out
for p in 1 2; do PSG_UMMA_PAIRS=$p timeout 300 python tools/bench_gemm.py fprop 2>&1 | tee gpurun_out/fprop_pairs$p.txt; done
