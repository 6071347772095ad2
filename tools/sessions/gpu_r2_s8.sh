set -x
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -q -m gpu -x 2>&1 | tail -25 > gpurun_out/r2s8_pytest.txt; tail -5 gpurun_out/r2s8_pytest.txt
timeout 600 python tools/variant_sweep.py --chunks 32 --only default,no_graph,default_again > gpurun_out/r2s8_sweep.json 2> gpurun_out/r2s8_sweep.err
grep -E "^[a-z0-9_]+/" gpurun_out/r2s8_sweep.err | cut -c1-330
VDL2GPU_SWEEP_TIMELINE=1 timeout 300 python tools/variant_sweep.py --chunks 32 --only default > gpurun_out/r2s8_timeline.json 2> gpurun_out/r2s8_timeline.err
timeout 300 python tools/block_trace.py interleaved > gpurun_out/r2s8_blocktrace.txt 2> gpurun_out/r2s8_bt.err; head -6 gpurun_out/r2s8_blocktrace.txt
timeout 600 python bench.py --no-repeat > gpurun_out/r2s8_bench.json 2> gpurun_out/r2s8_bench.err; tail -c 300 gpurun_out/r2s8_bench.err; cut -c1-300 gpurun_out/r2s8_bench.json
