#!/bin/bash
O=gpurun_out
python scripts/trace_update.py rainbow_atari_b1024 bf16 > $O/r07i_trace_rainbow_b1024_bf16.txt 2>&1
for lib in "" build/ab/libzoo_m5.so; do
  echo "== lib=${lib:-current}"
  ZOO_SM100_LIB=$lib python scripts/prof_update.py rainbow_atari_b1024 bf16 2>&1 | grep -E "launch_conv|TOTAL"
  for wl in rainbow_atari_b1024 iqn_atari_b512; do
    ZOO_SM100_LIB=$lib python bench.py --workload $wl --precision bf16 --no-cpu --no-c5 --steps 50 --warmup 5 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('$wl bf16 ms', d['ms_per_step'])"
  done
done
python bench.py --precision bf16 --no-cpu --no-c5 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('dqn b32 bf16 ms', d['ms_per_step'])"
python scripts/one_update.py rainbow_atari_b1024 bf16 1 3 > /dev/null 2>&1 && timeout 200 ncu --set full --clock-control none --import-source on --profile-from-start off -k regex:"stage_batch" -o $O/r07i_stage -f python scripts/one_update.py rainbow_atari_b1024 bf16 1 3 > $O/r07i_ncu.log 2>&1
