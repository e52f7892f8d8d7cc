#!/bin/bash
# Kernel-level and model-level GPU parity tests, in separate processes so one CUDA fault does not hide the rest.
cd "${GRAFT_REPO_ROOT:-.}"
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_kernels_gpu.py -q -m gpu -k "gemm" -p no:cacheprovider > gpurun_out/t_gemm.log 2>&1; echo "gemm rc=$?"
timeout 300 python -m pytest tests/test_kernels_gpu.py -q -m gpu -k "dwconv" -p no:cacheprovider > gpurun_out/t_dw.log 2>&1; echo "dw rc=$?"
timeout 300 python -m pytest tests/test_kernels_gpu.py -q -m gpu -k "convnext_block" -p no:cacheprovider > gpurun_out/t_blk.log 2>&1; echo "blk rc=$?"
timeout 300 python -m pytest tests/test_kernels_gpu.py -q -m gpu -k "not gemm and not dwconv and not convnext_block" -p no:cacheprovider > gpurun_out/t_rest.log 2>&1; echo "rest rc=$?"
timeout 900 python -m pytest tests/test_ga_convnext_model.py -q -m gpu -p no:cacheprovider > gpurun_out/t_model.log 2>&1; echo "model rc=$?"
timeout 600 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; echo "smoke rc=$?"
for f in t_gemm t_dw t_blk t_rest t_model smoke; do echo "== $f"; tail -4 gpurun_out/$f.log; done
timeout 900 python -m pytest tests/test_map_convnext_model.py -q -m gpu -p no:cacheprovider > gpurun_out/t_map.log 2>&1; echo "map rc=$?"
echo "== t_map"; tail -4 gpurun_out/t_map.log
