cd /root/repo
timeout 2400 python -m pytest tests -x -q -m gpu 2>&1 | tail -3
timeout 900 python bench.py --steps 10 --warmup 3 --no-infer --no-cpu-baseline --sustained 0 > gpurun_out/bench16.log 2>&1
echo rc=$?
python -c "import json; d=json.loads([l for l in open('gpurun_out/bench16.log') if l.startswith('{')][-1]); print(d['value'], d['ms_per_step'], d['gpu_launches'], d['e2e']['value'])"
timeout 600 python scripts/profile_step.py --top 400 --out gpurun_out/step_profile_h.txt --sequence gpurun_out/step_sequence_h.txt > /dev/null 2>gpurun_out/prof_err.txt
grep "gemm_tc2\|layernorm_bwd_kernel<float" gpurun_out/step_profile_h.txt | cut -c1-110
