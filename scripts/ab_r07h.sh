#!/bin/bash
for lib in "" build/ab/libzoo_m3.so build/ab/libzoo_m4.so; do
  echo "== lib=${lib:-current}"
  ZOO_SM100_LIB=$lib python scripts/prof_update.py rainbow_atari_b1024 bf16 2>&1 | grep -E "launch_conv_tma|TOTAL"
  ZOO_SM100_LIB=$lib python scripts/prof_update.py iqn_atari_b512 bf16 2>&1 | grep -E "launch_conv_tma|TOTAL"
  for wl in rainbow_atari_b1024 iqn_atari_b512; do
    ZOO_SM100_LIB=$lib python bench.py --workload $wl --precision bf16 --no-cpu --no-c5 --steps 50 --warmup 5 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('$wl bf16 ms', d['ms_per_step'])"
  done
done
