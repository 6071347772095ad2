# round 2, GPU session 3: K2 on 128-thread blocks (pipeline overlap must become deterministic), lean hypot, K3 warp RS + octet unstuffer, K1 alignment fix, streams
set -x
mkdir -p gpurun_out
timeout 1800 python -m pytest tests -q -m gpu -x 2>&1 | tail -25 > gpurun_out/r2s3_pytest.txt; tail -5 gpurun_out/r2s3_pytest.txt
timeout 900 python tools/variant_sweep.py > gpurun_out/r2s3_sweep.json 2> gpurun_out/r2s3_sweep.err; grep -E "^[a-z0-9_]+/" gpurun_out/r2s3_sweep.err | cut -c1-330
timeout 600 python tools/bench_streams.py --chunks 6 > gpurun_out/r2s3_streams.json 2> gpurun_out/r2s3_streams.err; tail -3 gpurun_out/r2s3_streams.err; head -60 gpurun_out/r2s3_streams.json
timeout 900 python bench.py > gpurun_out/r2s3_bench.json 2> gpurun_out/r2s3_bench.err; tail -c 400 gpurun_out/r2s3_bench.err; cut -c1-400 gpurun_out/r2s3_bench.json
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k2_sync -s 2 -c 1 -o gpurun_out/r2s3_k2 -f python tools/profile_run.py --chunks 4 > /dev/null 2>&1
timeout 600 ncu --set full --clock-control none -k regex:k2a_ -s 2 -c 1 -o gpurun_out/r2s3_k2a -f python tools/profile_run.py --chunks 4 > /dev/null 2>&1
timeout 600 ncu --set full --clock-control none -k regex:k3_burst -s 2 -c 1 -o gpurun_out/r2s3_k3 -f python tools/profile_run.py --chunks 4 > /dev/null 2>&1
timeout 900 ncu --set full --clock-control none -k regex:k1_mix_iir_decimate_lanes -s 1 -c 1 -o gpurun_out/r2s3_k1lanes -f python tools/bench_streams.py --chunks 1 > /dev/null 2>&1
ls -la gpurun_out/ | grep r2s3
