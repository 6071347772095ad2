# closing run with the ring walk as default (VDL2GPU_K2_VARIANT=3); variant 2 beside it for the A/B
timeout 300 python -m pytest tests -x -q -m gpu 2>&1 | tail -3
show() { python - "$1" <<'PY'
import json,sys
d=json.loads(open(sys.argv[1]).read())
print(sys.argv[1],'value',round(d['value']),'rt_ch',round(d['channels_at_realtime']),'e2e',round(d['e2e']['value']),'ms/step',round(d['ms_per_step'],2),{k:round(v,2) for k,v in d['roofline']['kernel_ms_per_launch'].items()}, d['parity']['pool_overflows'], d['parity']['out_overflows'], d['clocks']['sm_mhz'])
PY
}
timeout 200 python bench.py > gpurun_out/bench_close3.json 2> gpurun_out/bench_close3.err; tail -2 gpurun_out/bench_close3.err; show gpurun_out/bench_close3.json
VDL2GPU_K2_VARIANT=2 timeout 200 python bench.py --no-cpu-baseline > gpurun_out/bench_v2.json 2> gpurun_out/bench_v2.err; show gpurun_out/bench_v2.json
timeout 200 ncu --metrics gpu__time_duration.sum --clock-control none -s 16 -c 32 --csv --log-file gpurun_out/launches_final3.csv python tools/profile_run.py --chunks 8 > gpurun_out/launches_final3.log 2>&1; tail -1 gpurun_out/launches_final3.log | cut -c1-200
timeout 200 ncu --set full --clock-control none --import-source on -k regex:k2_sync -s 2 -c 1 -o gpurun_out/k2_final3 -f python tools/profile_run.py --chunks 4 > /dev/null 2>&1
timeout 200 ncu --metrics gpu__time_duration.sum --clock-control none -s 40 -c 120 --csv --log-file gpurun_out/launches_bench3.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/launches_bench3.log 2>&1
