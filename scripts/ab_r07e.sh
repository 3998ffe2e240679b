#!/bin/bash
O=gpurun_out
python -m pytest tests/test_gpu_gemm.py tests/test_gpu_full_size.py -m gpu -x -q -k "persistent_engine_all or (pinned and iqn)" > $O/r07e_tests.log 2>&1
tail -3 $O/r07e_tests.log
for v in 0 1 2 3; do
  echo "== ZOO_MERGE_BWD_VAR=$v"
  ZOO_MERGE_BWD_VAR=$v python scripts/prof_update.py iqn_atari_b512 bf16 2>&1 | grep -E "launch_persist|merge_bwd|TOTAL"
done
python bench.py --workload iqn_atari_b512 --precision bf16 --no-cpu --no-c5 --steps 50 --warmup 5 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('iqn bf16', d['ms_per_step'])"
