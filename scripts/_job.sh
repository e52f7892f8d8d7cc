cd /root/repo
timeout 1200 python bench.py > gpurun_out/bench_final.log 2> gpurun_out/bench_final.err; echo bench rc=$?
timeout 600 python bench.py --model map_convnext_tiny --batch 512 --no-infer --no-cpu-baseline --sustained 0 > gpurun_out/bench_map.log 2>&1; echo map rc=$?
timeout 600 python bench.py --model ga_CSWin_64_12211_tiny_224 --batch 128 --no-infer --no-cpu-baseline --sustained 0 > gpurun_out/bench_cswin.log 2>&1; echo cswin rc=$?
timeout 600 python scripts/profile_step.py --top 400 --out gpurun_out/step_profile_e.txt --sequence gpurun_out/step_sequence_e.txt > /dev/null 2>gpurun_out/prof_err.txt
head -2 gpurun_out/step_profile_e.txt
timeout 600 python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-e2e --no-graph --no-infer --sustained 0 > gpurun_out/b_plain.log 2>&1 || exit 1
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 16000 --csv --log-file gpurun_out/r02_launches.csv python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-e2e --no-graph --no-infer --sustained 0 > gpurun_out/ncu_c.log 2>&1
ls -la gpurun_out/r02_launches.csv
for f in bench_final bench_map bench_cswin; do grep '^{' gpurun_out/$f.log | tail -1 | python -c "import json,sys; d=json.loads(sys.stdin.read()); print(d['value'], d['ms_per_step'], d['gpu_launches'], d['e2e']['value'], d.get('infer'), d.get('sustained'))"; done
