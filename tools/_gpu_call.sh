set -x
mkdir -p gpurun_out
export BYZAGG_K2_SPLIT=1
( timeout 150 compute-sanitizer --tool memcheck --error-exitcode 9 python tools/prof_rules.py 25 5 70001 krum,bulyan 2>&1 | tail -4 ) > gpurun_out/r2_split_memcheck.txt 2>&1
( timeout 150 compute-sanitizer --tool racecheck --error-exitcode 9 python tools/prof_rules.py 23 5 33000 krum 2>&1 | tail -4 ) > gpurun_out/r2_split_racecheck.txt 2>&1
( timeout 700 python -m pytest tests -q -m gpu -x 2>&1 | tail -8 ) > gpurun_out/r2_t23.log 2>&1
unset BYZAGG_K2_SPLIT
( timeout 600 python tools/abbench.py byzantinemomentum_b200/libbyzagg.so@BYZAGG_K2_SPLIT=0 byzantinemomentum_b200/libbyzagg.so@BYZAGG_K2_SPLIT=1 --split 2>&1 | tail -10 ) > gpurun_out/r2_ab_split.txt
cat gpurun_out/r2_split_memcheck.txt gpurun_out/r2_split_racecheck.txt; tail -5 gpurun_out/r2_t23.log; cat gpurun_out/r2_ab_split.txt
