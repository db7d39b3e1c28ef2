mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/pytest.log
timeout 300 python tools/sweep_gn.py pf > gpurun_out/gn_sweep_pf.txt 2>&1; echo "sweep rc=$?"
PSG_GN_ONLY_CLUSTER=1 timeout 600 ncu --set full --import-source on --clock-control none --profile-from-start off -f -o gpurun_out/gn_729_320 python tools/ncu_gn.py 729 320 > gpurun_out/ncu_gn1.log 2>&1; echo "ncu1 rc=$?"
PSG_GN_ONLY_CLUSTER=1 timeout 600 ncu --set full --import-source on --clock-control none --profile-from-start off -f -o gpurun_out/gn_196_1280 python tools/ncu_gn.py 196 1280 > gpurun_out/ncu_gn2.log 2>&1; echo "ncu2 rc=$?"
ls -la gpurun_out/*.ncu-rep
