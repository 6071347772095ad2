set -x
mkdir -p gpurun_out
for v in "VDL2GPU_K2_VARIANT=4" "VDL2GPU_K2A_SPLIT=1" "VDL2GPU_K2A=0" "VDL2GPU_STAGES=3" "X=1"; do
  echo "== $v"
  env $v timeout 600 python -m pytest tests/test_gpu_parity.py -q -m gpu -x -k "frames_match or replicas" 2>&1 | tail -4
done > gpurun_out/r2s13_diag.txt 2>&1
cat gpurun_out/r2s13_diag.txt
