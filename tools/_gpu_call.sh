set -x
mkdir -p gpurun_out
( timeout 1200 python -m pytest tests -q -m gpu 2>&1 | tail -15 ) > gpurun_out/r2_t3.log 2>&1
( BYZAGG_K2_LEGACY=1 timeout 300 python tools/abk2.py 2>&1 | tail -12 ) > gpurun_out/r2_rules_legacy.log 2>&1
( BYZAGG_K2_NOFUSE=1 timeout 300 python tools/abk2.py 2>&1 | tail -12 ) > gpurun_out/r2_rules_nofuse.log 2>&1
( timeout 300 python tools/abk2.py 2>&1 | tail -12 ) > gpurun_out/r2_rules_fused.log 2>&1
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file gpurun_out/r2_rules_launches.csv python tools/prof_rules.py 25 5 1310922 krum,bulyan,cge,aksel > gpurun_out/ncu3.log 2>&1
tail -6 gpurun_out/r2_t3.log; cat gpurun_out/r2_rules_legacy.log gpurun_out/r2_rules_nofuse.log gpurun_out/r2_rules_fused.log
