# round 2: validation + evidence session for the current build
set -x
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -q -m gpu -x 2>&1 | tail -25 > gpurun_out/r2s12_pytest.txt; tail -4 gpurun_out/r2s12_pytest.txt
timeout 600 python tools/variant_sweep.py --chunks 32 --only default,no_graph,k2_ring_cpasync,default_again > gpurun_out/r2s12_sweep.json 2> gpurun_out/r2s12_sweep.err
grep -E "^[a-z0-9_]+/" gpurun_out/r2s12_sweep.err | cut -c1-330
timeout 900 python bench.py > gpurun_out/r2s12_bench.json 2> gpurun_out/r2s12_bench.err; tail -c 300 gpurun_out/r2s12_bench.err; cut -c1-300 gpurun_out/r2s12_bench.json
timeout 600 python bench.py --impl reference --steps 6 --warmup 3 > gpurun_out/r2s12_bench_reference.json 2>/dev/null; cut -c1-300 gpurun_out/r2s12_bench_reference.json
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r2s12_launches_bench.csv python bench.py --steps 1 --warmup 3 --chunks-per-step 8 --no-repeat --no-parity --no-cpu-baseline > gpurun_out/r2s12_launches_bench.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k1_mix_iir_decimate_packed -s 2 -c 1 -o gpurun_out/r2s12_k1 -f python tools/profile_run.py --chunks 4 > /dev/null 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k2_sync -s 2 -c 1 -o gpurun_out/r2s12_k2 -f python tools/profile_run.py --chunks 4 > /dev/null 2>&1
timeout 600 ncu --set full --clock-control none -k regex:k2a_ -s 2 -c 1 -o gpurun_out/r2s12_k2a -f python tools/profile_run.py --chunks 4 > /dev/null 2>&1
timeout 600 ncu --set full --clock-control none -k regex:k3_burst -s 2 -c 1 -o gpurun_out/r2s12_k3 -f python tools/profile_run.py --chunks 4 > /dev/null 2>&1
ls -la gpurun_out | grep r2s12
