# round 2: K2a with loads one step ahead + straight-line atan2: parity + timing
set -x
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -q -m gpu -x 2>&1 | tail -25 > gpurun_out/r2s17_pytest.txt; tail -4 gpurun_out/r2s17_pytest.txt
timeout 600 python tools/variant_sweep.py --chunks 32 --only default,default_again > gpurun_out/r2s17_sweep.json 2> gpurun_out/r2s17_sweep.err
grep -E "^[a-z0-9_]+/" gpurun_out/r2s17_sweep.err | cut -c1-330
timeout 600 ncu --set full --clock-control none -k regex:k2a_ -s 2 -c 1 -o gpurun_out/r2s17_k2a -f python tools/profile_run.py --chunks 4 > /dev/null 2>&1
ls -la gpurun_out | grep r2s17
