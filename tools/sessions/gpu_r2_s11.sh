set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_stage.py -q -m gpu -x -k "stream" 2>&1 | tail -8 > gpurun_out/r2s11_pytest.txt; tail -3 gpurun_out/r2s11_pytest.txt
timeout 600 python tools/bench_streams.py --chunks 6 > gpurun_out/r2s11_streams.json 2> gpurun_out/r2s11_streams.err; tail -3 gpurun_out/r2s11_streams.err; head -40 gpurun_out/r2s11_streams.json
