set -x
mkdir -p gpurun_out
( timeout 900 python -m pytest tests/test_cuda_study_step.py tests/test_attack_py_gpu.py -q -m gpu -x 2>&1 | tail -30 ) > gpurun_out/r2_t17.log 2>&1
( timeout 300 python tools/study_time.py 2>&1 | tail -6 ) > gpurun_out/r2_study_time.txt
( timeout 300 python tools/study_time.py 25 36489290 2 2>&1 | tail -6 ) >> gpurun_out/r2_study_time.txt
tail -12 gpurun_out/r2_t17.log; cat gpurun_out/r2_study_time.txt
