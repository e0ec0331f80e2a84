#!/bin/bash
# Full single-GPU validation of the current tree (run through gpurun): GPU tests, bench, ncu.
# Usage: tools/gpu_validate.sh <tag>
tag=${1:-x}
mkdir -p gpurun_out
timeout 500 python -m pytest tests -m gpu -q -x 2>&1 | tail -3
timeout 300 python bench.py --steps 200 --warmup 10 > gpurun_out/bench_$tag.json 2> gpurun_out/bench_$tag.err
tail -3 gpurun_out/bench_$tag.err
python - <<PY
import json
l = json.load(open("gpurun_out/bench_$tag.json"))
s = l.pop("sweep", [])
print(json.dumps(l)[:2800])
for r in s:
  if "ms" in r: print("%-8s n=%2d f=%2d d=%9d %9.1f us frac %.3f ro %.3f" % (r["gar"], r["n"], r["f"], r["d"], r["ms"]*1e3, r["frac"], r["read_only_frac"]))
  else: print(r)
PY
timeout 200 ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file gpurun_out/launches_$tag.csv python bench.py --steps 5 --warmup 3 --no-sweep > gpurun_out/ncu_b_$tag.log 2>&1
timeout 200 ncu --set full --clock-control none --import-source on -k regex:k1_sorted -s 3 -c 2 -o gpurun_out/prof_trmean_$tag python bench.py --steps 5 --warmup 3 --no-sweep > gpurun_out/ncu_c_$tag.log 2>&1
ls gpurun_out | tail -4
timeout 200 ncu --set full --clock-control none --import-source on -k regex:k6_study -c 2 -o gpurun_out/prof_study_$tag python - > gpurun_out/ncu_d_$tag.log 2>&1 <<PY
import torch, byzantinemomentum_b200 as bz
rows = [torch.randn(36489290, device="cuda") for _ in range(25)]
for _ in range(2):
  bz.engine.avg_dev_max_async(rows)
torch.cuda.synchronize()
PY
ls gpurun_out | tail -4
