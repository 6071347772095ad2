# round-1 closing evidence on one B200: config-4 AWGN sweep, refreshed launch lists + K3 capture, gpu tests, default bench
set -x
timeout 600 python tools/awgn_sweep.py --out gpurun_out/awgn_sweep.json > gpurun_out/awgn_sweep.txt 2> gpurun_out/awgn_sweep.err; tail -18 gpurun_out/awgn_sweep.txt; tail -2 gpurun_out/awgn_sweep.err | cut -c1-400
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -s 16 -c 32 --csv --log-file gpurun_out/launches_final.csv python tools/profile_run.py --chunks 8 > gpurun_out/launches_final.log 2>&1; tail -1 gpurun_out/launches_final.log | cut -c1-300
timeout 300 ncu --set full --clock-control none --import-source on -k regex:k3_burst -s 2 -c 1 -o gpurun_out/k3_final -f python tools/profile_run.py --chunks 4 > /dev/null 2>&1
timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none -s 40 -c 120 --csv --log-file gpurun_out/launches_bench.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/launches_bench.log 2>&1; tail -c 300 gpurun_out/launches_bench.log
timeout 900 python -m pytest tests -x -q -m gpu 2>&1 | tail -3
timeout 600 python bench.py > gpurun_out/bench_close.json 2> gpurun_out/bench_close.err; tail -2 gpurun_out/bench_close.err; cut -c1-1500 gpurun_out/bench_close.json
