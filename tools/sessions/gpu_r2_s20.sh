# round 2: branch-layout hints in the walk: parity + A/B timing
set -x
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -q -m gpu -x 2>&1 | tail -25 > gpurun_out/r2s20_pytest.txt; tail -4 gpurun_out/r2s20_pytest.txt
timeout 600 python tools/variant_sweep.py --chunks 32 --only default,default_again > gpurun_out/r2s20_sweep.json 2> gpurun_out/r2s20_sweep.err
grep -E "^[a-z0-9_]+/" gpurun_out/r2s20_sweep.err | cut -c1-330
ls -la gpurun_out | grep r2s20
