set -x
mkdir -p gpurun_out
timeout 1800 python -m pytest tests -q -m gpu -x 2>&1 | tail -25 > gpurun_out/r2s6_pytest.txt; tail -5 gpurun_out/r2s6_pytest.txt
timeout 300 python tools/block_trace.py interleaved > gpurun_out/r2s6_blocktrace_interleaved.txt 2> gpurun_out/r2s6_bt.err; tail -2 gpurun_out/r2s6_bt.err; head -8 gpurun_out/r2s6_blocktrace_interleaved.txt
timeout 900 python tools/variant_sweep.py --chunks 32 --only default,no_graph,k2_plane,k2_ring_cpasync,default_again,no_graph_again > gpurun_out/r2s6_sweep.json 2> gpurun_out/r2s6_sweep.err
grep -E "^[a-z0-9_]+/" gpurun_out/r2s6_sweep.err | cut -c1-330
VDL2GPU_BALANCE=0 timeout 900 python tools/variant_sweep.py --chunks 32 --only default,default_again > gpurun_out/r2s6_sweep_nobalance.json 2> gpurun_out/r2s6_sweep_nobalance.err
grep -E "^[a-z0-9_]+/" gpurun_out/r2s6_sweep_nobalance.err | cut -c1-330
timeout 900 python bench.py > gpurun_out/r2s6_bench.json 2> gpurun_out/r2s6_bench.err; tail -c 400 gpurun_out/r2s6_bench.err; cut -c1-400 gpurun_out/r2s6_bench.json
