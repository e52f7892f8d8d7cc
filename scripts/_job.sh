cd /root/repo
timeout 2400 python -m pytest tests -x -q -m gpu 2>&1 | tail -4
timeout 900 python bench.py --steps 10 --warmup 3 --no-infer --no-cpu-baseline --sustained 0 2>gpurun_out/bench19.err | python -c "import json,sys; d=json.loads([l for l in sys.stdin if l.startswith('{')][-1]); print(d['value'], d['ms_per_step'], d['gpu_launches'], d['e2e']['value'])"; tail -3 gpurun_out/bench19.err
