set -x
timeout 1200 python -m pytest tests -x -q -m gpu -s 2>&1 | grep -E "passed|failed|Es/N0|Error|error" | tail -12
timeout 600 compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "cfg2 and frames" 2>&1 | tail -6
