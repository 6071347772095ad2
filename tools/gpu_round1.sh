set -x
python __graft_entry__.py --smoke 2>&1 | tail -2
python bench.py --steps 4 --warmup 3 > gpurun_out/bench_r01_a.json 2> gpurun_out/bench_r01_a.err; tail -c 3000 gpurun_out/bench_r01_a.json; tail -5 gpurun_out/bench_r01_a.err
python bench.py --steps 3 --warmup 3 --k1-scalar --no-cpu-baseline > gpurun_out/bench_r01_scalar.json 2>&1; tail -c 1500 gpurun_out/bench_r01_scalar.json
python bench.py --steps 3 --warmup 3 --channels 65536 --no-cpu-baseline > gpurun_out/bench_r01_65536.json 2>&1; tail -c 1500 gpurun_out/bench_r01_65536.json
ncu --metrics gpu__time_duration.sum --clock-control none -c 40 --csv --log-file gpurun_out/launches_r01.csv python tools/profile_run.py --chunks 6 > gpurun_out/launches_r01.log 2>&1; tail -3 gpurun_out/launches_r01.log
ncu --set full --clock-control none --import-source on -k regex:k1_mix -s 2 -c 1 -o gpurun_out/k1_r01 -f python tools/profile_run.py --chunks 4 > gpurun_out/ncu_k1.log 2>&1; tail -3 gpurun_out/ncu_k1.log
ncu --set full --clock-control none --import-source on -k regex:k2_sync -s 2 -c 1 -o gpurun_out/k2_r01 -f python tools/profile_run.py --chunks 4 > gpurun_out/ncu_k2.log 2>&1; tail -3 gpurun_out/ncu_k2.log
