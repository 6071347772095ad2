# round 2, GPU session 2: K2a fix, 4-warp K1 with the conflict-free NCO table, fully staged K2 walk, one-stream-per-channel HBM mode
set -x
mkdir -p gpurun_out
timeout 1800 python -m pytest tests -q -m gpu -x 2>&1 | tail -25 > gpurun_out/r2s2_pytest.txt; tail -5 gpurun_out/r2s2_pytest.txt
timeout 900 python tools/variant_sweep.py > gpurun_out/r2s2_sweep.json 2> gpurun_out/r2s2_sweep.err; grep -E "^[a-z0-9_]+/" gpurun_out/r2s2_sweep.err | cut -c1-400
VDL2GPU_SWEEP_TIMELINE=1 timeout 600 python tools/variant_sweep.py --only default,no_graph > gpurun_out/r2s2_sweep_timeline.json 2> gpurun_out/r2s2_sweep_timeline.err
timeout 600 python tools/bench_streams.py --chunks 6 > gpurun_out/r2s2_streams.json 2> gpurun_out/r2s2_streams.err; tail -3 gpurun_out/r2s2_streams.err; cat gpurun_out/r2s2_streams.json | head -50
timeout 900 python bench.py > gpurun_out/r2s2_bench.json 2> gpurun_out/r2s2_bench.err; tail -c 400 gpurun_out/r2s2_bench.err; cut -c1-400 gpurun_out/r2s2_bench.json
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k1_mix_iir_decimate_packed -s 2 -c 1 -o gpurun_out/r2s2_k1 -f python tools/profile_run.py --chunks 4 > /dev/null 2>&1
VDL2GPU_K2_VARIANT=5 timeout 600 ncu --set full --clock-control none --import-source on -k regex:k2_sync -s 2 -c 1 -o gpurun_out/r2s2_k2v5 -f python tools/profile_run.py --chunks 4 > /dev/null 2>&1
timeout 600 ncu --set full --clock-control none -k regex:k2a_ -s 2 -c 1 -o gpurun_out/r2s2_k2a -f python tools/profile_run.py --chunks 4 > /dev/null 2>&1
timeout 900 ncu --set full --clock-control none -k regex:k1_mix_iir_decimate_lanes -s 1 -c 1 -o gpurun_out/r2s2_k1lanes -f python tools/bench_streams.py --chunks 1 > /dev/null 2>&1
ls -la gpurun_out/ | grep r2s2
