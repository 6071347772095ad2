timeout 900 python -m pytest tests -x -q -m gpu 2>&1 | tail -4
ncu --set full --clock-control none --import-source on -k regex:k2_sync -s 2 -c 1 -o gpurun_out/k2_r01b -f python tools/profile_run.py --chunks 4 > gpurun_out/ncu_k2b.log 2>&1; tail -2 gpurun_out/ncu_k2b.log
ncu --set full --clock-control none --import-source on -k regex:k2a_ -s 2 -c 1 -o gpurun_out/k2a_r01b -f python tools/profile_run.py --chunks 4 > gpurun_out/ncu_k2ab.log 2>&1; tail -2 gpurun_out/ncu_k2ab.log
