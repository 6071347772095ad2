set -x
timeout 600 python -m pytest tests -x -q -m gpu 2>&1 | tail -4
for v in 0 2; do VDL2GPU_K1_VARIANT=$v python bench.py --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/bench_v$v.json 2>gpurun_out/bench_v$v.err; python - <<PY
import json
d=json.load(open('gpurun_out/bench_v$v.json'))
print('variant $v value',round(d['value']),'rt_ch',round(d['channels_at_realtime']),'e2e',round(d['e2e']['value']),'ms/step',round(d['ms_per_step'],2),d['roofline']['kernel_ms_per_launch'],'frac',round(d['roofline']['frac'],3))
PY
done
