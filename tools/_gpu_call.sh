set -x
mkdir -p gpurun_out
( timeout 300 python bench.py --impl reference --steps 5 --warmup 1 2> gpurun_out/r2_bench7.err | tail -1 ) > gpurun_out/r2_bench7_ref.json
( timeout 900 python bench.py 2>> gpurun_out/r2_bench7.err | tail -1 ) > gpurun_out/r2_bench7.json
timeout 300 ncu --set full --clock-control none --import-source on -k regex:k2_ring -s 2 -c 1 -o gpurun_out/r2_prof_k2_n51_w16 python tools/prof_rules.py 51 12 4568373 krum > gpurun_out/r2_ncu_k2w16.log 2>&1
ncu -i gpurun_out/r2_prof_k2_n51_w16.ncu-rep --page raw --csv > gpurun_out/r2_k2ring_n51_w16_ncu_raw.csv 2>/dev/null
tail -3 gpurun_out/r2_bench7.err; tail -3 gpurun_out/r2_ncu_k2w16.log
python - <<'PY'
import json, csv
l=json.loads(open('gpurun_out/r2_bench7.json').read())
print({k:l[k] for k in ('value','ms_per_step','gpu_launches','clocks')}); print(l['roofline']['frac'], l['e2e']['ms_per_step'], l['e2e']['host_path']); print(l['cpu_baseline']['ms_per_call'], l['cpu_baseline']['kind'])
r=json.loads(open('gpurun_out/r2_bench7_ref.json').read()); print(r['value'], r['ms_per_step'], r['config']==l['config'])
for row in l.get('sweep',[]):
  if row.get('gar') in ('krum','bulyan') and row.get('d') in (1310922, 4568373): print({k:(round(v,4) if isinstance(v,float) else v) for k,v in row.items()})
rows=list(csv.reader(open('gpurun_out/r2_k2ring_n51_w16_ncu_raw.csv')))
h=rows[0]; r=rows[2]
for i,x in enumerate(h):
  if x in ("Kernel Name","gpu__time_duration.sum","launch__cluster_size","launch__cluster_max_active","launch__grid_size","launch__registers_per_thread","sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active","sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_elapsed","smsp__issue_active.avg.pct_of_peak_sustained_active","sm__warps_active.avg.pct_of_peak_sustained_active","dram__bytes_read.sum"): print(x, r[i], rows[1][i])
PY
