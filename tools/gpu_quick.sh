timeout 900 python -m pytest tests -x -q -m gpu 2>&1 | tail -4
for kv in 1; do
VDL2GPU_K2_VARIANT=$kv python bench.py --steps 4 --warmup 3 --no-cpu-baseline > gpurun_out/bench_q$kv.json 2>gpurun_out/bench_q$kv.err; tail -3 gpurun_out/bench_q$kv.err; python - <<PY
import json
d=json.load(open('gpurun_out/bench_q$kv.json'))
print('K2 variant $kv: value',round(d['value']),'rt_ch',round(d['channels_at_realtime']),'e2e',round(d['e2e']['value']),'ms/step',round(d['ms_per_step'],2),{k:round(v,2) for k,v in d['roofline']['kernel_ms_per_launch'].items()},'frac',round(d['roofline']['frac'],3), d['clocks']['sm_mhz'], d['parity'])
PY
done
