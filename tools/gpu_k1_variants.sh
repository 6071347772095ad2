timeout 600 python -m pytest tests/test_gpu_api.py -x -q -m gpu -k "planar" 2>&1 | tail -2
for v in 2 5 6 7; do VDL2GPU_K1_VARIANT=$v python tools/profile_run.py --chunks 12 2>&1 | grep -o "kernel ms {[^}]*}" ; done
