mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_kernels.py -m gpu -x -q -k "groupnorm" > gpurun_out/pytest_gn.log 2>&1; echo "pytest rc=$?"; tail -5 gpurun_out/pytest_gn.log
timeout 200 python tools/bench_gn.py 256 fast > gpurun_out/gn_stream_default.txt 2>&1; echo "rc=$?"; cat gpurun_out/gn_stream_default.txt
for gb in 8 16 48 96; do echo "group MB $gb"; PSG_GNS_TUNE="$((gb*1048576)),0,0" timeout 200 python tools/bench_gn.py 256 stream 2>&1 | tee gpurun_out/gn_stream_g$gb.txt; done
for rows in 32 96 128; do echo "rows $rows"; PSG_GNS_TUNE="25165824,$rows,0" timeout 200 python tools/bench_gn.py 256 stream 2>&1 | tee gpurun_out/gn_stream_r$rows.txt; done
