#!/bin/bash
for lib in "" build/ab/libzoo_s6.so build/ab/libzoo_s8.so; do
  echo "== lib=${lib:-current}"
  ZOO_SM100_LIB=$lib python scripts/prof_update.py rainbow_atari_b1024 bf16 2>&1 | grep -E "stage_batch|TOTAL"
  ZOO_SM100_LIB=$lib python scripts/prof_update.py dqn_atari_b32 fp32 2>&1 | grep -E "stage_batch|TOTAL"
  ZOO_SM100_LIB=$lib python scripts/prof_update.py iqn_atari_b512 fp32 2>&1 | grep -E "stage_batch|merge_bwd|TOTAL"
done
python -m pytest tests/test_gpu_learners.py -m gpu -x -q -k "pipeline or atari" 2>&1 | tail -3
