ncu --metrics gpu__time_duration.sum --clock-control none -s 40 -c 120 --csv --log-file gpurun_out/launches_bench.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/launches_bench.log 2>&1; tail -c 300 gpurun_out/launches_bench.log
for ch in 256 4096 65536; do python bench.py --steps 4 --warmup 3 --no-cpu-baseline --channels $ch > gpurun_out/bench_ch$ch.json 2>gpurun_out/bench_ch$ch.err; python - <<PY
import json
d=json.loads(open("gpurun_out/bench_ch$ch.json").read())
print($ch,"value",round(d["value"]),"rt",round(d["channels_at_realtime"]),"ms/step",round(d["ms_per_step"],2),"e2e",round(d["e2e"]["value"]),{k:round(v,3) for k,v in d["roofline"]["kernel_ms_per_launch"].items()},"frac",round(d["roofline"]["frac"],3), d["parity"]["pool_overflows"], d["parity"]["out_overflows"])
PY
done
