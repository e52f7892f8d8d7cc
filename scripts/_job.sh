cd /root/repo
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; echo smoke rc=$?; tail -1 gpurun_out/smoke.log
timeout 1200 python scripts/parity_report.py gpurun_out/parity_report.json > gpurun_out/parity_report.log 2>&1; echo parity rc=$?
timeout 1200 python bench.py > gpurun_out/bench_final.log 2> gpurun_out/bench_final.err; echo bench rc=$?
timeout 1200 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_ref.log 2> gpurun_out/bench_ref.err; echo ref rc=$?
timeout 600 python bench.py --model map_convnext_tiny --batch 512 --no-infer --no-cpu-baseline --sustained 0 > gpurun_out/bench_map.log 2>&1; echo map rc=$?
timeout 600 python bench.py --model ga_CSWin_64_12211_tiny_224 --batch 128 --no-infer --no-cpu-baseline --sustained 0 > gpurun_out/bench_cswin.log 2>&1; echo cswin rc=$?
timeout 600 python scripts/profile_step.py --top 400 --out gpurun_out/step_profile_final.txt --sequence gpurun_out/step_sequence_final.txt > /dev/null 2>gpurun_out/prof_err.txt
head -2 gpurun_out/step_profile_final.txt
for f in bench_final bench_map bench_cswin; do grep '^{' gpurun_out/$f.log | tail -1 | python -c "import json,sys; d=json.loads(sys.stdin.read()); print(d['value'], d['ms_per_step'], d['gpu_launches'], d['e2e']['value'], d.get('infer'), d.get('sustained'), d['roofline']['kernel'][:60], d['roofline']['frac'], d['roofline']['traffic'])"; done
