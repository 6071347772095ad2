# round 2, 2-GPU session: hardware parity on every rank for both fan-out modes, weak + strong bench lines at N=2, launch list with the fan-out active
set -x
mkdir -p gpurun_out
nvidia-smi -L; nvidia-smi topo -m | head -8
timeout 900 python -m pytest tests/test_gpu_multi.py -q -m gpu -x 2>&1 | tail -15 > gpurun_out/r2n2_pytest.txt; tail -4 gpurun_out/r2n2_pytest.txt
for mode in ce nccl; do
  timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 3 --warmup 3 --chunks-per-step 128 --no-repeat --fanout $mode > gpurun_out/r2n2_bench_weak_$mode.json 2> gpurun_out/r2n2_bench_weak_$mode.err; tail -c 300 gpurun_out/r2n2_bench_weak_$mode.err; cut -c1-300 gpurun_out/r2n2_bench_weak_$mode.json
done
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 --steps 3 --warmup 3 --chunks-per-step 128 --no-repeat --scaling strong > gpurun_out/r2n2_bench_strong.json 2> gpurun_out/r2n2_bench_strong.err; cut -c1-300 gpurun_out/r2n2_bench_strong.json
timeout 900 python bench.py --steps 3 --warmup 3 --chunks-per-step 128 --no-repeat > gpurun_out/r2n2_bench_n1.json 2> gpurun_out/r2n2_bench_n1.err; cut -c1-300 gpurun_out/r2n2_bench_n1.json
ls -la gpurun_out | grep r2n2
