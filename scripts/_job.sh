cd /root/repo
timeout 900 python -m pytest tests/test_kernels_gpu.py -x -q -m gpu -k "stacked_linear" 2>&1 | grep -E "^E|passed|failed|Error" | head -20
