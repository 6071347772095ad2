# experiment: K2 unwrap through the transition table (VDL2GPU_K2_VARIANT=2) vs TwoSum (1); channel counts that fill all 592 sub-partitions
VDL2GPU_K2_VARIANT=2 timeout 600 python -m pytest tests/test_gpu_parity.py tests/test_gpu_scale.py -x -q 2>&1 | tail -2
show() { python - "$1" <<'PY'
import json,sys
d=json.loads(open(sys.argv[1]).read())
print(sys.argv[1],'value',round(d['value']),'rt_ch',round(d['channels_at_realtime']),'e2e',round(d['e2e']['value']),'ms/step',round(d['ms_per_step'],2),{k:round(v,2) for k,v in d['roofline']['kernel_ms_per_launch'].items()}, d['parity']['pool_overflows'], d['parity']['out_overflows'])
PY
}
VDL2GPU_K2_VARIANT=1 python bench.py --no-cpu-baseline > gpurun_out/exp_v1.json 2>gpurun_out/exp_v1.err; show gpurun_out/exp_v1.json
VDL2GPU_K2_VARIANT=2 python bench.py --no-cpu-baseline > gpurun_out/exp_v2.json 2>gpurun_out/exp_v2.err; show gpurun_out/exp_v2.json
VDL2GPU_K2_VARIANT=2 python bench.py --no-cpu-baseline --channels 18944 > gpurun_out/exp_v2_18944.json 2>gpurun_out/exp_v2_18944.err; show gpurun_out/exp_v2_18944.json
VDL2GPU_K2_VARIANT=1 python bench.py --no-cpu-baseline --channels 18944 > gpurun_out/exp_v1_18944.json 2>gpurun_out/exp_v1_18944.err; show gpurun_out/exp_v1_18944.json
VDL2GPU_K2_VARIANT=2 python bench.py --no-cpu-baseline --channels 37888 --steps 4 > gpurun_out/exp_v2_37888.json 2>gpurun_out/exp_v2_37888.err; show gpurun_out/exp_v2_37888.json
