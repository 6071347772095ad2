# round 2: validation + evidence session after the walk-tail fix (magnitudes computed in the walk)
set -x
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -q -m gpu -x 2>&1 | tail -25 > gpurun_out/r2s14_pytest.txt; tail -4 gpurun_out/r2s14_pytest.txt
timeout 900 python bench.py > gpurun_out/r2s14_bench.json 2> gpurun_out/r2s14_bench.err; tail -c 300 gpurun_out/r2s14_bench.err; cut -c1-400 gpurun_out/r2s14_bench.json
timeout 300 python tools/bench_streams.py --chunks 8 > gpurun_out/r2s14_streams.json 2> gpurun_out/r2s14_streams.err; cut -c1-400 gpurun_out/r2s14_streams.json
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r2s14_launches_bench.csv python bench.py --steps 1 --warmup 3 --chunks-per-step 8 --no-repeat --no-parity --no-cpu-baseline > gpurun_out/r2s14_launches_bench.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k2_sync -s 2 -c 1 -o gpurun_out/r2s14_k2 -f python tools/profile_run.py --chunks 4 > /dev/null 2>&1
timeout 600 ncu --set full --clock-control none -k regex:k2a_ -s 2 -c 1 -o gpurun_out/r2s14_k2a -f python tools/profile_run.py --chunks 4 > /dev/null 2>&1
timeout 600 ncu --set full --clock-control none -k regex:k1_mix_iir_decimate_lanes -s 1 -c 1 -o gpurun_out/r2s14_k1lanes -f python tools/bench_streams.py --chunks 2 > /dev/null 2>&1
ls -la gpurun_out | grep r2s14
