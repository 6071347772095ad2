# last GPU seconds of the round: 18944 channels (every warp full) against 16384 on the same box, short legs
set -x
mkdir -p gpurun_out
timeout 80 python bench.py --channels 18944 --steps 2 --warmup 3 --chunks-per-step 64 --no-repeat --no-cpu-baseline > gpurun_out/r2last_18944.json 2> gpurun_out/r2last_18944.err; echo "rc=$?"; cut -c1-200 gpurun_out/r2last_18944.json
timeout 60 python bench.py --channels 16384 --steps 2 --warmup 3 --chunks-per-step 64 --no-repeat --no-cpu-baseline --no-parity > gpurun_out/r2last_16384.json 2> gpurun_out/r2last_16384.err; echo "rc=$?"; cut -c1-200 gpurun_out/r2last_16384.json
