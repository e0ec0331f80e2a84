set -x
mkdir -p gpurun_out
( timeout 300 python tools/e2e_parts.py 2>&1 | tail -12 ) > gpurun_out/r2_e2e_parts.log 2>&1
( timeout 600 python bench.py --steps 200 --warmup 10 --no-sweep 2> gpurun_out/r2_bench4.err | tail -1 ) > gpurun_out/r2_bench4.json
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -s 125 -c 30 --csv --log-file gpurun_out/r2_bench_launches.csv python bench.py --steps 5 --warmup 3 --no-sweep > gpurun_out/ncu4.log 2>&1
cat gpurun_out/r2_e2e_parts.log; tail -3 gpurun_out/r2_bench4.err; python -c "
import json
l=json.loads(open('gpurun_out/r2_bench4.json').read())
print(l['value'], l['ms_per_step'], l['roofline']['frac']); print(json.dumps(l['e2e'],indent=1))"
