# round 2, second 2-GPU session (final build): hardware parity on every rank for both fan-out modes, weak + strong lines at N=2,
# copy-engine fan-out under a strict timeout, and the opt-in fused-phase K1 through the variant test
set -x
mkdir -p gpurun_out
timeout 400 python -m pytest tests/test_gpu_multi.py -q -m gpu -x 2>&1 | tail -15 > gpurun_out/r2n2b_pytest.txt; tail -4 gpurun_out/r2n2b_pytest.txt
timeout 500 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 6 --warmup 3 --no-repeat > gpurun_out/r2n2b_bench_weak_nccl.json 2> gpurun_out/r2n2b_bench_weak_nccl.err; tail -c 300 gpurun_out/r2n2b_bench_weak_nccl.err; cut -c1-300 gpurun_out/r2n2b_bench_weak_nccl.json
timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 --steps 6 --warmup 3 --no-repeat --scaling strong > gpurun_out/r2n2b_bench_strong_nccl.json 2> gpurun_out/r2n2b_bench_strong_nccl.err; tail -c 300 gpurun_out/r2n2b_bench_strong_nccl.err; cut -c1-300 gpurun_out/r2n2b_bench_strong_nccl.json
timeout 240 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29513 bench.py --gpus 2 --steps 3 --warmup 3 --chunks-per-step 64 --no-repeat --fanout ce > gpurun_out/r2n2b_bench_weak_ce.json 2> gpurun_out/r2n2b_bench_weak_ce.err; echo "ce rc=$?"; tail -c 300 gpurun_out/r2n2b_bench_weak_ce.err; cut -c1-300 gpurun_out/r2n2b_bench_weak_ce.json
timeout 300 python -m pytest tests/test_gpu_stage.py -q -m gpu -x -k "variants" 2>&1 | tail -5 > gpurun_out/r2n2b_pytest_variants.txt; tail -3 gpurun_out/r2n2b_pytest_variants.txt
ls -la gpurun_out | grep r2n2b
