# round 2: K1 with the two-operation NCO table address: parity, final bench line, launch list, K1 capture
set -x
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -q -m gpu -x 2>&1 | tail -25 > gpurun_out/r2s19_pytest.txt; tail -4 gpurun_out/r2s19_pytest.txt
timeout 900 python bench.py > gpurun_out/r2s19_bench.json 2> gpurun_out/r2s19_bench.err; tail -c 300 gpurun_out/r2s19_bench.err; cut -c1-300 gpurun_out/r2s19_bench.json
timeout 600 python tools/variant_sweep.py --chunks 32 --only default > gpurun_out/r2s19_sweep.json 2> gpurun_out/r2s19_sweep.err
grep -E "^[a-z0-9_]+/" gpurun_out/r2s19_sweep.err | cut -c1-330
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r2s19_launches_bench.csv python bench.py --steps 1 --warmup 3 --chunks-per-step 8 --no-repeat --no-parity --no-cpu-baseline > gpurun_out/r2s19_launches_bench.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k1_mix_iir_decimate_packed -s 2 -c 1 -o gpurun_out/r2s19_k1 -f python tools/profile_run.py --chunks 4 > /dev/null 2>&1
ls -la gpurun_out | grep r2s19
