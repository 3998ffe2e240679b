#!/bin/bash
# timing ablations of the persistent engine's epilogue (results are wrong under ablation): per-launch device times of the IQN 512 x 64 bf16 update
# ZOO_PERSIST_ABLATE bits: 1 = no TMA stores, 2 = no conversion / staging, 4 = old (unpacked) epilogue arithmetic
for a in 0 1 2 3 4; do
  echo "== ZOO_PERSIST_ABLATE=$a"
  ZOO_PERSIST_ABLATE=$a python scripts/prof_update.py iqn_atari_b512 bf16 2>&1 | grep -E "launch_persist|TOTAL"
done
