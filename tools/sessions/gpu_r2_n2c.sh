# round 2, last 2-GPU session (final build): weak and strong lines with in-run parity on both ranks
set -x
mkdir -p gpurun_out
timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29531 bench.py --gpus 2 --steps 4 --warmup 3 --chunks-per-step 128 --no-repeat > gpurun_out/r2n2c_bench_weak_nccl.json 2> gpurun_out/r2n2c_bench_weak_nccl.err; echo "rc=$?"; cut -c1-200 gpurun_out/r2n2c_bench_weak_nccl.json
timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29532 bench.py --gpus 2 --steps 4 --warmup 3 --chunks-per-step 128 --no-repeat --scaling strong > gpurun_out/r2n2c_bench_strong_nccl.json 2> gpurun_out/r2n2c_bench_strong_nccl.err; echo "rc=$?"; cut -c1-200 gpurun_out/r2n2c_bench_strong_nccl.json
