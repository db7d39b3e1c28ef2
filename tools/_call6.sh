mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_kernels.py -m gpu -x -q -k "groupnorm" > gpurun_out/pytest_gn.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/pytest_gn.log
timeout 200 python tools/bench_gn.py 256 cluster 2>&1 | tee gpurun_out/gn_final.txt
timeout 400 python bench.py --steps 20 --warmup 5 > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "bench rc=$?"
python - <<'P'
import json
d=json.loads(open('gpurun_out/bench.json').read().strip().splitlines()[-1])
print(d['ms_per_step'], d['value'], d['e2e']['value'], json.dumps(d['hbm_kernels'])[:600], d['cpu_baseline'])
P
