set -x
mkdir -p gpurun_out
( timeout 1200 python -m pytest tests -q -m gpu -x 2>&1 | tail -8 ) > gpurun_out/r2_t11.log 2>&1
( timeout 300 python tools/k2_ab.py --cases 11:1310922,16:1310922,20:1310922,11:36489290,5:1310922,2:1310922 --json gpurun_out/k2_ab_list.json 2>&1 | tail -12 ) > gpurun_out/r2_k2ab_list.log 2>&1
( BYZAGG_K2_NOLIST=1 timeout 300 python tools/k2_ab.py --cases 11:1310922,16:1310922,20:1310922 --no-alias --only ring 2>&1 | tail -5 ) > gpurun_out/r2_k2ab_nolist.log 2>&1
( timeout 300 python tools/abk2.py 2>&1 | tail -8 ) > gpurun_out/r2_rules6.log 2>&1
tail -4 gpurun_out/r2_t11.log; cat gpurun_out/r2_k2ab_list.log gpurun_out/r2_k2ab_nolist.log gpurun_out/r2_rules6.log
