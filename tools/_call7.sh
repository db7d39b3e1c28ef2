mkdir -p gpurun_out
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 10 --warmup 3 > gpurun_out/bench_2gpu.json 2> gpurun_out/bench_2gpu.err; echo "bench2 rc=$?"
cut -c1-300 gpurun_out/bench_2gpu.json; tail -3 gpurun_out/bench_2gpu.err
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 tools/check_dp_equivalence.py > gpurun_out/dp_equiv.txt 2>&1; echo "equiv rc=$?"; tail -8 gpurun_out/dp_equiv.txt
timeout 300 python bench.py --gpus 1 --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/bench_1gpu_samebox.json 2>/dev/null; cut -c1-200 gpurun_out/bench_1gpu_samebox.json
