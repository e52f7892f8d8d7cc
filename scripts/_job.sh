cd /root/repo
timeout 1200 python -m pytest tests/test_parity_baseline_shapes.py tests/test_ga_convnext_model.py tests/test_ga_cswin.py tests/test_engine_gpu.py -x -q -m gpu 2>&1 | tail -8
echo "== per-branch"; GA_BATCH_HEADS=0 timeout 900 python bench.py --steps 10 --warmup 3 --no-infer --no-cpu-baseline --sustained 0 2>/dev/null | python -c "import json,sys; d=json.loads([l for l in sys.stdin if l.startswith('{')][-1]); print(d['value'], d['ms_per_step'], d['gpu_launches'])"
echo "== batched"; timeout 900 python bench.py --steps 10 --warmup 3 --no-infer --no-cpu-baseline --sustained 0 2>gpurun_out/bench17.err | python -c "import json,sys; d=json.loads([l for l in sys.stdin if l.startswith('{')][-1]); print(d['value'], d['ms_per_step'], d['gpu_launches'])"; tail -3 gpurun_out/bench17.err
