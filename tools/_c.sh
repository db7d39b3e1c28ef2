mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_full.log 2>&1; echo "pytest rc=$?"; tail -2 gpurun_out/pytest_full.log
timeout 300 python tools/host_profile.py 256 2>&1 | head -3
timeout 400 python bench.py --steps 20 --warmup 5 > gpurun_out/bench_final.json 2> gpurun_out/bench_final.err; echo "bench rc=$?"
python -c "
import json
d=json.loads(open('gpurun_out/bench_final.json').read().strip().splitlines()[-1])
print(d['ms_per_step'], d['value'], d['e2e']['value'], d['roofline']['achieved'], d['roofline']['frac'], d['roofline']['traffic'], d['clocks'], d['gpu_launches'])
"
