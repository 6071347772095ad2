timeout 900 python -m pytest tests -x -q -m gpu 2>&1 | tail -4
python bench.py --steps 4 --warmup 3 --no-cpu-baseline > gpurun_out/bench_q.json 2>gpurun_out/bench_q.err; tail -3 gpurun_out/bench_q.err; python - <<PY
import json
d=json.load(open('gpurun_out/bench_q.json'))
print('value',round(d['value']),'rt_ch',round(d['channels_at_realtime']),'e2e',round(d['e2e']['value']),'ms/step',round(d['ms_per_step'],2),d['roofline']['kernel_ms_per_launch'],'frac',round(d['roofline']['frac'],3), d['clocks'])
PY
