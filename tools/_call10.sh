mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gemm_engines.py tests/test_unet_b256.py tests/test_unet_parity.py tests/test_trainer.py -m gpu -x -q > gpurun_out/pytest_gemm.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/pytest_gemm.log
for i in 1 2; do
timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/bench_tnpairs1_$i.json 2>/dev/null; python -c "import json;d=json.loads(open('gpurun_out/bench_tnpairs1_$i.json').read().splitlines()[-1]);print('tn_pairs on ',d['ms_per_step'],d['roofline']['achieved'])"
PSG_TN_PAIRS=0 timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/bench_tnpairs0_$i.json 2>/dev/null; python -c "import json;d=json.loads(open('gpurun_out/bench_tnpairs0_$i.json').read().splitlines()[-1]);print('tn_pairs off',d['ms_per_step'],d['roofline']['achieved'])"
done
