# BASELINE config 3: 256 channels fanned out from one stream; ncu on K1 (DRAM bytes, issue, FMA pipe) + launch list
timeout 300 ncu --set full --clock-control none --import-source on -k regex:k1_mix -s 2 -c 1 -o gpurun_out/k1_256ch -f python tools/profile_run.py --chunks 4 --channels 256 > gpurun_out/k1_256ch.log 2>&1; tail -1 gpurun_out/k1_256ch.log | cut -c1-200
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -s 16 -c 32 --csv --log-file gpurun_out/launches_256ch.csv python tools/profile_run.py --chunks 8 --channels 256 > gpurun_out/launches_256ch.log 2>&1
timeout 300 python bench.py --channels 256 --no-cpu-baseline > gpurun_out/bench_256ch.json 2>gpurun_out/bench_256ch.err; cut -c1-400 gpurun_out/bench_256ch.json
