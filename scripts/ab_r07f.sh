#!/bin/bash
for k in 1024 512 64; do
  echo "== ZOO_PAIR_MIN_K=$k"
  ZOO_PAIR_MIN_K=$k python scripts/prof_update.py iqn_atari_b512 bf16 2>&1 | grep -E "launch_persist|merge_bwd|TOTAL"
done
python scripts/prof_update.py rainbow_atari_b1024 bf16 2>&1
