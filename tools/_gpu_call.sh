set -x
mkdir -p gpurun_out
L=byzantinemomentum_b200
( timeout 600 python tools/abbench.py --dist $L/libbyzagg.so $L/libbyzagg-notrig.so 2>&1 | tail -8 ) > gpurun_out/r2_ab_trigger.log 2>&1
( timeout 200 python tools/k2_ab.py --cases 25:1310922,25:36489290 --no-alias --only ring 2>&1 | tail -3 ) > gpurun_out/r2_k2_now.log 2>&1
( BYZAGG_LIBRARY=$PWD/$L/libbyzagg-notrig.so timeout 200 python tools/k2_ab.py --cases 25:1310922,25:36489290 --no-alias --only ring 2>&1 | tail -3 ) > gpurun_out/r2_k2_notrig.log 2>&1
nvidia-smi --query-gpu=clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active --format=csv > gpurun_out/r2_clocks.txt
cat gpurun_out/r2_ab_trigger.log gpurun_out/r2_k2_now.log gpurun_out/r2_k2_notrig.log gpurun_out/r2_clocks.txt
