#!/bin/bash
O=gpurun_out
python -m pytest tests/test_gpu_learners.py tests/test_gpu_full_size.py -m gpu -x -q -k "rainbow or qr or rem or c51 or dueling" > $O/r07l_tests.log 2>&1
tail -3 $O/r07l_tests.log
for h in 1 0; do
  echo "== ZOO_HEAD_BF16=$h"
  ZOO_HEAD_BF16=$h python scripts/prof_update.py rainbow_atari_b1024 bf16 2>&1 | grep -E "t8|stage_batch|TOTAL"
  ZOO_HEAD_BF16=$h python bench.py --workload rainbow_atari_b1024 --precision bf16 --no-cpu --no-c5 --steps 50 --warmup 5 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('rainbow bf16 ms', d['ms_per_step'])"
done
python scripts/full_size_parity.py 2>&1 | grep -A4 "rainbow_atari_b1024  bf16"
