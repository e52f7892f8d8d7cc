cd /root/repo
timeout 900 python -m pytest tests/test_kernels_gpu.py -x -q -m gpu -k "fused_layernorm or gemm or block" 2>&1 | tail -6
echo "== unfused"; GA_FUSE_LN_BWD=0 timeout 900 python bench.py --steps 10 --warmup 3 --no-infer --no-cpu-baseline --sustained 0 2>/dev/null | python -c "import json,sys; d=json.loads([l for l in sys.stdin if l.startswith('{')][-1]); print(d['value'], d['ms_per_step'], d['gpu_launches'])"
echo "== fused"; timeout 900 python bench.py --steps 10 --warmup 3 --no-infer --no-cpu-baseline --sustained 0 2>/dev/null | python -c "import json,sys; d=json.loads([l for l in sys.stdin if l.startswith('{')][-1]); print(d['value'], d['ms_per_step'], d['gpu_launches'])"
