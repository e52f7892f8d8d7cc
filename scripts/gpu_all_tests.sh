#!/bin/bash
# every GPU suite, one process per group
cd "${GRAFT_REPO_ROOT:-.}"
bash scripts/gpu_kernel_tests.sh 2>&1 | grep -E "rc=|passed|failed|error" 
TAILN=3 bash scripts/gpu_cswin_tests.sh 2>&1 | grep -E "rc=|passed|failed|error"
timeout 900 python -m pytest tests/test_engine_gpu.py -q -m gpu -p no:cacheprovider 2>&1 | tail -5
