mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1"
timeout 300 $TR --master-port 29541 tools/check_dp_equivalence.py 2>&1 | grep dp_equivalence
P='import json,sys;d=json.loads(open(sys.argv[1]).read().splitlines()[-1]);print(sys.argv[1],d["ms_per_step"],d["value"])'
timeout 300 $TR --master-port 29542 bench.py --gpus 2 --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/b2_side1.json 2>/dev/null; python -c "$P" gpurun_out/b2_side1.json
PSG_BUCKET_ISSUE=main timeout 300 $TR --master-port 29543 bench.py --gpus 2 --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/b2_main1.json 2>/dev/null; python -c "$P" gpurun_out/b2_main1.json
timeout 300 $TR --master-port 29544 bench.py --gpus 2 --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/b2_side2.json 2>/dev/null; python -c "$P" gpurun_out/b2_side2.json
PSG_BUCKET_ISSUE=main timeout 300 $TR --master-port 29545 bench.py --gpus 2 --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/b2_main2.json 2>/dev/null; python -c "$P" gpurun_out/b2_main2.json
