set -x
mkdir -p gpurun_out
( timeout 900 python -m pytest tests/test_cuda_host_pipeline.py tests/test_cuda_abi.py tests/test_cuda_golden.py tests/test_cuda_vs_reference_fullsize.py -q -m gpu -x -k "bulyan or pipeline or host or golden or Bulyan" 2>&1 | tail -15 ) > gpurun_out/r2_t18.log 2>&1
( timeout 600 python tools/abbench.py byzantinemomentum_b200/libbyzagg-base.so byzantinemomentum_b200/libbyzagg.so --bulyan 2>&1 | tail -12 ) > gpurun_out/r2_ab_k4.txt
( timeout 300 python tools/e2e_pipeline_ab.py 2>&1 | tail -12 ) > gpurun_out/r2_e2e_pipeline_ab.txt
tail -6 gpurun_out/r2_t18.log; cat gpurun_out/r2_ab_k4.txt gpurun_out/r2_e2e_pipeline_ab.txt
