# round 2, 4-GPU session (final build): weak-scaling line with in-run parity on every rank
set -x
mkdir -p gpurun_out
timeout 420 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29521 bench.py --gpus 4 --steps 4 --warmup 3 --chunks-per-step 128 --no-repeat > gpurun_out/r2n4_bench_weak_nccl.json 2> gpurun_out/r2n4_bench_weak_nccl.err; echo "rc=$?"; tail -c 400 gpurun_out/r2n4_bench_weak_nccl.err; cut -c1-300 gpurun_out/r2n4_bench_weak_nccl.json
ls -la gpurun_out | grep r2n4
