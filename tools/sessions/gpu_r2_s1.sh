# round 2, GPU session 1: parity of everything new, the new bench line, variant A/B, launch list + ncu of the three big kernels
set -x
mkdir -p gpurun_out
nvidia-smi -L
timeout 1800 python -m pytest tests -q -m gpu 2>&1 | tail -25 > gpurun_out/r2s1_pytest.txt; tail -5 gpurun_out/r2s1_pytest.txt
timeout 900 python bench.py > gpurun_out/r2s1_bench.json 2> gpurun_out/r2s1_bench.err; tail -c 600 gpurun_out/r2s1_bench.err; cut -c1-1500 gpurun_out/r2s1_bench.json
timeout 900 python tools/variant_sweep.py > gpurun_out/r2s1_sweep.json 2> gpurun_out/r2s1_sweep.err; tail -20 gpurun_out/r2s1_sweep.err
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -s 16 -c 40 --csv --log-file gpurun_out/r2s1_launches.csv python tools/profile_run.py --chunks 8 > gpurun_out/r2s1_launches.log 2>&1; tail -1 gpurun_out/r2s1_launches.log | cut -c1-300
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k1_mix -s 2 -c 1 -o gpurun_out/r2s1_k1 -f python tools/profile_run.py --chunks 4 > /dev/null 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k2_sync -s 2 -c 1 -o gpurun_out/r2s1_k2 -f python tools/profile_run.py --chunks 4 > /dev/null 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k2a_ -s 2 -c 1 -o gpurun_out/r2s1_k2a -f python tools/profile_run.py --chunks 4 > /dev/null 2>&1
ls -la gpurun_out/ | tail -20
