for v in 2 5; do VDL2GPU_K1_VARIANT=$v python tools/profile_run.py --chunks 12 2>&1 | grep -o "kernel ms {[^}]*}" ; done
VDL2GPU_K1_VARIANT=5 timeout 600 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "k1_dec" 2>&1 | tail -2
