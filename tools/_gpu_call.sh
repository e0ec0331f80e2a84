set -x
mkdir -p gpurun_out
( timeout 1500 python -m pytest tests -q -m gpu 2>&1 | tail -8 ) > gpurun_out/r2_t13.log 2>&1
( timeout 900 python bench.py --steps 200 --warmup 10 2> gpurun_out/r2_bench2.err | tail -1 ) > gpurun_out/r2_bench2.json
( timeout 300 python bench.py --impl reference --steps 5 --warmup 1 2>> gpurun_out/r2_bench2.err | tail -1 ) > gpurun_out/r2_bench2_ref.json
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -s 30 -c 40 --csv --log-file gpurun_out/r2_bench_launches.csv python bench.py --steps 5 --warmup 3 --no-sweep > gpurun_out/ncu4.log 2>&1
tail -4 gpurun_out/r2_t13.log; tail -3 gpurun_out/r2_bench2.err; python - <<'PY'
import json
l=json.loads(open('gpurun_out/r2_bench2.json').read())
print({k:l[k] for k in ('value','ms_per_step','gpu_launches','clocks')}); print(l['roofline']); print(l['e2e']); print(l['cpu_baseline']); print(l['config'])
r=json.loads(open('gpurun_out/r2_bench2_ref.json').read()); print(r['value'], r['ms_per_step'], r['cpu_baseline'], r['config']==l['config'])
for row in l.get('sweep',[])[:12]: print({k:(round(v,4) if isinstance(v,float) else v) for k,v in row.items()})
PY
