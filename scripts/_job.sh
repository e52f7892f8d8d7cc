cd /root/repo
timeout 600 python -m pytest tests/test_kernels_gpu.py tests/test_engine_gpu.py -x -q -m gpu -k "arena or prep or engine or graph" 2>&1 | tail -6
timeout 900 python bench.py --steps 10 --warmup 3 --no-infer --no-cpu-baseline --sustained 0 > gpurun_out/bench10.log 2>&1
echo rc=$?
python -c "import json; d=json.loads([l for l in open('gpurun_out/bench10.log') if l.startswith('{')][-1]); print(d['value'], d['ms_per_step'], d['gpu_launches'], d['e2e']['value'])"
tail -3 gpurun_out/bench10.log | cut -c1-300
