timeout 1200 python -m pytest tests -x -q -m gpu 2>&1 | tail -3
python bench.py --no-cpu-baseline > gpurun_out/bench_q1.json 2>gpurun_out/bench_q1.err; tail -2 gpurun_out/bench_q1.err; python - <<PY
import json
d=json.loads(open('gpurun_out/bench_q1.json').read())
print('value',round(d['value']),'rt_ch',round(d['channels_at_realtime']),'e2e',round(d['e2e']['value']),'ms/step',round(d['ms_per_step'],2),{k:round(v,2) for k,v in d['roofline']['kernel_ms_per_launch'].items()},'frac',round(d['roofline']['frac'],3), d['clocks']['sm_mhz'], d['parity'])
PY
