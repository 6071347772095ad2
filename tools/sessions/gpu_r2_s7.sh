set -x
mkdir -p gpurun_out
timeout 900 python tools/variant_sweep.py --chunks 32 --only default,k2a_overlap,k2a_overlap_libm,default_again > gpurun_out/r2s7_sweep.json 2> gpurun_out/r2s7_sweep.err
grep -E "^[a-z0-9_]+/" gpurun_out/r2s7_sweep.err | cut -c1-330
VDL2GPU_K2A_EXCLUSIVE=0 timeout 300 python tools/block_trace.py interleaved > gpurun_out/r2s7_blocktrace_overlap.txt 2> gpurun_out/r2s7_bt.err; head -6 gpurun_out/r2s7_blocktrace_overlap.txt
VDL2GPU_SWEEP_TIMELINE=1 timeout 600 python tools/variant_sweep.py --chunks 32 --only default,k2a_overlap > gpurun_out/r2s7_timeline.json 2> gpurun_out/r2s7_timeline.err
