set -x
mkdir -p gpurun_out
timeout 900 python tools/variant_sweep.py --chunks 32 --only split8,split16,split32,split64,default_again > gpurun_out/r2s10_sweep.json 2> gpurun_out/r2s10_sweep.err
grep -E "^[a-z0-9_]+/" gpurun_out/r2s10_sweep.err | cut -c1-330
