set -x
mkdir -p gpurun_out
( timeout 1500 python -m pytest tests -q -m gpu -x 2>&1 | tail -8 ) > gpurun_out/r2_t22.log 2>&1
( timeout 900 python tools/abbench.py byzantinemomentum_b200/libbyzagg.so@BYZAGG_K2_W16=0 byzantinemomentum_b200/libbyzagg.so@BYZAGG_K2_W16=1 --wide 2>&1 | tail -12 ) > gpurun_out/r2_ab_w16.txt
tail -5 gpurun_out/r2_t22.log; cat gpurun_out/r2_ab_w16.txt
