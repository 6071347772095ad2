# round 2: fused phase pass in K1 (no K2a launch): parity + A/B timing
set -x
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -q -m gpu -x 2>&1 | tail -25 > gpurun_out/r2s15_pytest.txt; tail -4 gpurun_out/r2s15_pytest.txt
VDL2GPU_SWEEP_TIMELINE=1 timeout 600 python tools/variant_sweep.py --chunks 32 --only default,k2a_separate,default_again > gpurun_out/r2s15_sweep.json 2> gpurun_out/r2s15_sweep.err
grep -E "^[a-z0-9_]+/" gpurun_out/r2s15_sweep.err | cut -c1-330
timeout 300 python tools/block_trace.py > gpurun_out/r2s15_blocktrace.txt 2>&1; tail -12 gpurun_out/r2s15_blocktrace.txt | cut -c1-250
ls -la gpurun_out | grep r2s15
