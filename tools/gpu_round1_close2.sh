# closing run after the table-unwrap walker: full gpu suite, default bench (with CPU baseline), refreshed launch lists and K2 capture
timeout 900 python -m pytest tests -x -q -m gpu 2>&1 | tail -3
timeout 600 python bench.py > gpurun_out/bench_close.json 2> gpurun_out/bench_close.err; tail -2 gpurun_out/bench_close.err; cut -c1-300 gpurun_out/bench_close.json
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -s 16 -c 32 --csv --log-file gpurun_out/launches_final.csv python tools/profile_run.py --chunks 8 > gpurun_out/launches_final.log 2>&1; tail -1 gpurun_out/launches_final.log | cut -c1-200
timeout 300 ncu --set full --clock-control none --import-source on -k regex:k2_sync -s 2 -c 1 -o gpurun_out/k2_final -f python tools/profile_run.py --chunks 4 > /dev/null 2>&1
timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none -s 40 -c 120 --csv --log-file gpurun_out/launches_bench.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/launches_bench.log 2>&1; tail -c 200 gpurun_out/launches_bench.log
timeout 300 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_ref.json 2> gpurun_out/bench_ref.err; cut -c1-400 gpurun_out/bench_ref.json
