set -x
mkdir -p gpurun_out
( BYZAGG_H2D_BATCH=1 timeout 300 python tools/e2e_pipeline_ab.py 2>&1 | tail -12 ) > gpurun_out/r2_e2e_pipeline_ab_batched.txt
( BYZAGG_H2D_BATCH=0 timeout 300 python tools/e2e_pipeline_ab.py 2>&1 | tail -12 ) > gpurun_out/r2_e2e_pipeline_ab_loop.txt
( timeout 300 python -m pytest tests/test_cuda_host_pipeline.py -q -m gpu -x 2>&1 | tail -5 ) > gpurun_out/r2_t19.log 2>&1
cat gpurun_out/r2_e2e_pipeline_ab_batched.txt gpurun_out/r2_e2e_pipeline_ab_loop.txt gpurun_out/r2_t19.log
