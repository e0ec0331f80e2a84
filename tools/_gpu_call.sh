set -x
mkdir -p gpurun_out
( timeout 900 python -m pytest tests/test_cuda_reuse.py tests/test_cuda_gradient_rows.py tests/test_attack_py_gpu.py -q -m gpu 2>&1 | tail -30 ) > gpurun_out/r2_t9.log 2>&1
( timeout 300 python tools/linesearch_time.py gpurun_out/r2_linesearch.json 2>&1 | tail -8 ) > gpurun_out/r2_linesearch.log 2>&1
( timeout 900 python bench.py --steps 200 --warmup 10 2> gpurun_out/r2_bench1.err | tail -1 ) > gpurun_out/r2_bench1.json
tail -30 gpurun_out/r2_t9.log; cat gpurun_out/r2_linesearch.log; tail -3 gpurun_out/r2_bench1.err; python - <<'PY'
import json
l=json.loads(open('gpurun_out/r2_bench1.json').read())
print({k:l[k] for k in ('value','ms_per_step','gpu_launches','clocks')}); print(l['roofline']); print(l['e2e']); print(l['cpu_baseline'])
for r in l.get('sweep',[]): print(r)
PY
