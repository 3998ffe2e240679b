#!/bin/bash
O=gpurun_out
python -m pytest tests/test_gpu_gemm.py tests/test_gpu_full_size.py tests/test_gpu_learners.py -m gpu -x -q -k "persistent_engine_all or (pinned and iqn) or iqn" > $O/r07m_tests.log 2>&1
tail -3 $O/r07m_tests.log
for ew in 16 8; do
  echo "== ZOO_PERSIST_EW=$ew"
  ZOO_PERSIST_EW=$ew python scripts/prof_update.py iqn_atari_b512 bf16 2>&1 | grep -E "launch_persist|TOTAL"
  ZOO_PERSIST_EW=$ew python bench.py --workload iqn_atari_b512 --precision bf16 --no-cpu --no-c5 --steps 50 --warmup 5 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('iqn bf16 ms', d['ms_per_step'])"
done
