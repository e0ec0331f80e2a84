set -x
mkdir -p gpurun_out
( timeout 1200 python -m pytest tests -q -m gpu -x 2>&1 | tail -8 ) > gpurun_out/r2_t4.log 2>&1
( timeout 300 python tools/abk2.py 2>&1 | tail -12 ) > gpurun_out/r2_rules_fused2.log 2>&1
( BYZAGG_K2_BIGTILE=1 timeout 300 python tools/k2_ab.py --cases 40:1310922,51:1310922,51:4568373 --no-alias --only ring 2>&1 | tail -12 ) > gpurun_out/r2_bigtile.log 2>&1
( timeout 300 python tools/k2_ab.py --cases 25:1310922,40:1310922,51:1310922,51:4568373 --no-alias --only ring 2>&1 | tail -12 ) > gpurun_out/r2_smalltile.log 2>&1
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file gpurun_out/r2_rules_launches2.csv python tools/prof_rules.py 25 5 1310922 krum,bulyan > gpurun_out/ncu3.log 2>&1
tail -4 gpurun_out/r2_t4.log; cat gpurun_out/r2_rules_fused2.log gpurun_out/r2_bigtile.log gpurun_out/r2_smalltile.log
