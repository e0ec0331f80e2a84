set -x
mkdir -p gpurun_out
( timeout 1500 python -m pytest tests -q -m gpu 2>&1 | tail -8 ) > gpurun_out/r2_t20.log 2>&1
( timeout 200 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -3 ) > gpurun_out/r2_smoke2.log 2>&1
( timeout 300 python bench.py --impl reference --steps 5 --warmup 1 2> gpurun_out/r2_bench6.err | tail -1 ) > gpurun_out/r2_bench6_ref.json
( timeout 900 python bench.py 2>> gpurun_out/r2_bench6.err | tail -1 ) > gpurun_out/r2_bench6.json
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 300 --csv --log-file gpurun_out/r2_launches_bench_final.csv python bench.py --steps 5 --warmup 3 --no-sweep --no-sharded > gpurun_out/r2_ncu_b.log 2>&1
python tools/ncu_summary.py gpurun_out/r2_launches_bench_final.csv > gpurun_out/r2_launches_bench_final_summary.txt 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:k4_bulyan -s 2 -c 1 -o gpurun_out/r2_prof_k4 python tools/prof_rules.py 25 5 1310922 bulyan > gpurun_out/r2_ncu_k4.log 2>&1
ncu -i gpurun_out/r2_prof_k4.ncu-rep --page raw --csv > gpurun_out/r2_k4_ncu_raw.csv 2>/dev/null
tail -4 gpurun_out/r2_t20.log; cat gpurun_out/r2_smoke2.log; tail -3 gpurun_out/r2_bench6.err; cat gpurun_out/r2_launches_bench_final_summary.txt | head -20
python - <<'PY'
import json
l=json.loads(open('gpurun_out/r2_bench6.json').read())
print({k:l[k] for k in ('value','ms_per_step','gpu_launches','clocks')}); print(l['roofline']['frac'], l['e2e']['ms_per_step'], l['e2e']['host_path'], l['e2e']['h2d_probe']); print(l['cpu_baseline']['ms_per_call'], l['cpu_baseline']['kind'])
r=json.loads(open('gpurun_out/r2_bench6_ref.json').read()); print(r['value'], r['ms_per_step'], r['config']==l['config'])
for row in l.get('sweep',[]):
  if row.get('gar') in ('krum','bulyan') and row.get('d') in (1310922, 4568373): print({k:(round(v,4) if isinstance(v,float) else v) for k,v in row.items()})
PY
