#!/bin/bash
# Round-2 opening A/B (one gpurun call): build the variants HERE first (no GPU needed), then
#   gpurun --timeout 900 -- 'tools/ab_round2.sh run'
# Variants: t64 = K1 with 64-thread CTAs (tail of the 4.3-wave launch at d = 1.31M),
#           t1024 = K2 TMA tiles of 1024 columns, 2 stages (half the bulk copies per coordinate),
#           hint = mbarrier.try_wait with a 20 us suspend hint (the probe loop is 11 % of K2's instructions),
#           flush2 = K2 flushes its fp32 accumulators into fp64 every 32 terms instead of 16.
# (each variant .so is 61 MB: delete them after the run, they travel with every gpurun push)
set -e
cd "$(dirname "$0")/.."
if [ "$1" != "run" ]; then
  make -C byzantinemomentum_b200/csrc -j8
  make -C byzantinemomentum_b200/csrc -j8 VARIANT=t64 EXTRA=-DBZ_K1_THREADS=64
  make -C byzantinemomentum_b200/csrc -j8 VARIANT=t1024 EXTRA="-DBZ_K2_TMA_T=1024 -DBZ_K2_TMA_MIN_STAGES=2"
  make -C byzantinemomentum_b200/csrc -j8 VARIANT=hint EXTRA="-DBZ_MBAR_HINT_NS=20000"
  make -C byzantinemomentum_b200/csrc -j8 VARIANT=flush2 EXTRA="-DBZ_K2_FLUSH_SCALE=2"
  ls -la byzantinemomentum_b200/libbyzagg*.so
  exit 0
fi
L=byzantinemomentum_b200
timeout 300 python tools/abbench.py $L/libbyzagg.so $L/libbyzagg-t64.so
timeout 400 python tools/abbench.py --dist $L/libbyzagg.so $L/libbyzagg-t1024.so $L/libbyzagg-hint.so $L/libbyzagg-flush2.so
BYZAGG_LIBRARY=$PWD/$L/libbyzagg-t1024.so timeout 300 python -m pytest tests/test_cuda_fullsize.py tests/test_cuda_sharded_phases.py -m gpu -q -x 2>&1 | tail -2
BYZAGG_LIBRARY=$PWD/$L/libbyzagg-t64.so timeout 300 python -m pytest tests/test_cuda_abi.py tests/test_cuda_fullsize.py -m gpu -q -x 2>&1 | tail -2
