set -x
mkdir -p gpurun_out
VDL2GPU_SWEEP_TIMELINE=1 timeout 900 python tools/variant_sweep.py --chunks 32 --only default,k2_plane,k2_staged_1warp,k2_plane_1warp,all_1warp,k1_one_warp > gpurun_out/r2s4_timeline.json 2> gpurun_out/r2s4_timeline.err
grep -E "^[a-z0-9_]+/" gpurun_out/r2s4_timeline.err | cut -c1-200
timeout 900 python tools/variant_sweep.py --chunks 32 --only default,k2_plane,k2_staged_1warp,k2_plane_1warp,all_1warp,k1_one_warp > gpurun_out/r2s4_sweep.json 2> gpurun_out/r2s4_sweep.err
grep -E "^[a-z0-9_]+/" gpurun_out/r2s4_sweep.err | cut -c1-330
