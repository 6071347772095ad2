set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_scale.py tests/test_gpu_api.py -q -m gpu -x 2>&1 | tail -8 > gpurun_out/r2s9_pytest.txt; tail -3 gpurun_out/r2s9_pytest.txt
timeout 900 python tools/variant_sweep.py --chunks 32 --only split1,split2,split3,split4,default_again > gpurun_out/r2s9_sweep.json 2> gpurun_out/r2s9_sweep.err
grep -E "^[a-z0-9_]+/" gpurun_out/r2s9_sweep.err | cut -c1-330
VDL2GPU_SWEEP_TIMELINE=1 timeout 300 python tools/variant_sweep.py --chunks 32 --only default > gpurun_out/r2s9_timeline.json 2> gpurun_out/r2s9_timeline.err
