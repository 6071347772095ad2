set -x
mkdir -p gpurun_out
timeout 300 python tools/block_trace.py interleaved > gpurun_out/r2s5_blocktrace_interleaved.txt 2> gpurun_out/r2s5_bt.err; tail -3 gpurun_out/r2s5_bt.err; cat gpurun_out/r2s5_blocktrace_interleaved.txt
timeout 300 python tools/block_trace.py replica > gpurun_out/r2s5_blocktrace_replica.txt 2>> gpurun_out/r2s5_bt.err; cat gpurun_out/r2s5_blocktrace_replica.txt
VDL2GPU_K2_VARIANT=2 timeout 300 python tools/block_trace.py interleaved > gpurun_out/r2s5_blocktrace_plane.txt 2>> gpurun_out/r2s5_bt.err; cat gpurun_out/r2s5_blocktrace_plane.txt
