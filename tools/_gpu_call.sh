set -x
mkdir -p gpurun_out
( timeout 600 python -m pytest tests -q -m gpu -x 2>&1 | tail -4 ) > gpurun_out/r2_t24.log 2>&1
( timeout 200 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2 ) > gpurun_out/r2_smoke3.log 2>&1
( timeout 600 python bench.py 2> gpurun_out/r2_bench8.err | tail -1 ) > gpurun_out/r2_bench8.json
tail -3 gpurun_out/r2_t24.log; cat gpurun_out/r2_smoke3.log; tail -2 gpurun_out/r2_bench8.err
python - <<'PY'
import json
l=json.loads(open('gpurun_out/r2_bench8.json').read())
print({k:l[k] for k in ('value','ms_per_step','gpu_launches','clocks')}); e=l['e2e']; print(l['roofline']['frac'], e['ms_per_step'], e.get('ms_per_step_min'), e.get('ms_per_step_median'), e.get('ms_per_step_max'), e['host_path'])
for row in l.get('sweep',[]):
  if row.get('gar') in ('krum','bulyan') and row.get('d') in (1310922, 4568373): print({k:(round(v,4) if isinstance(v,float) else v) for k,v in row.items()})
PY
