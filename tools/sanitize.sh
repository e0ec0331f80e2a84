#!/bin/bash
# compute-sanitizer passes over the CUDA path (run through gpurun); logs under gpurun_out/.
mkdir -p gpurun_out
export PYTHONFAULTHANDLER=1
timeout 900 compute-sanitizer --tool memcheck --error-exitcode 7 --log-file gpurun_out/memcheck.log \
  python -m pytest tests/test_cuda_abi.py -m gpu -q -x -k "ragged or unaligned or error_codes or specialised or staged or (every_n and (n11 or n25 or n51 or n64 or n3 or n1))" > gpurun_out/memcheck_pytest.log 2>&1
echo "memcheck rc=$?"; tail -3 gpurun_out/memcheck_pytest.log; grep -E "ERROR SUMMARY|Invalid|out of bounds" gpurun_out/memcheck.log | head -5
timeout 600 compute-sanitizer --tool memcheck --error-exitcode 7 --log-file gpurun_out/memcheck_dist.log \
  python -m pytest tests/test_cuda_abi.py tests/test_cuda_sharded_phases.py -m gpu -q -x -k "distance_rules and (empire-11 or empire-25 or nan-26 or empire-51 or empire-64 or empire-5-) or phases_compose" > gpurun_out/memcheck_dist_pytest.log 2>&1
echo "memcheck(dist) rc=$?"; tail -3 gpurun_out/memcheck_dist_pytest.log; grep -E "ERROR SUMMARY|Invalid|out of bounds" gpurun_out/memcheck_dist.log | head -5
timeout 600 compute-sanitizer --tool racecheck --error-exitcode 7 --log-file gpurun_out/racecheck.log \
  python -m pytest tests/test_cuda_abi.py -m gpu -q -x -k "distance_rules and (empire-11 or empire-25-5 or empire-51)" > gpurun_out/racecheck_pytest.log 2>&1
echo "racecheck rc=$?"; tail -3 gpurun_out/racecheck_pytest.log; grep -E "RACECHECK SUMMARY|hazard" gpurun_out/racecheck.log | head -5
