#!/bin/bash
# GPU check of a change: the whole GPU test-suite, then the three bench lines that matter and their in-situ profiles.  usage: quick_r07.sh <tag> [pytest -k expr]
tag=${1:-r07}
O=gpurun_out
if [ -n "$2" ]; then python -m pytest tests -m gpu -x -q -k "$2" > $O/${tag}_tests.log 2>&1; else python -m pytest tests -m gpu -x -q > $O/${tag}_tests.log 2>&1; fi
tail -3 $O/${tag}_tests.log
python bench.py --no-cpu --no-c5 > $O/${tag}_bench_dqn_b32_fp32.json 2>/dev/null
python bench.py --workload iqn_atari_b512 --precision bf16 --no-cpu --no-c5 --steps 50 --warmup 5 > $O/${tag}_bench_iqn_b512_bf16.json 2>/dev/null
python bench.py --workload rainbow_atari_b1024 --precision bf16 --no-cpu --no-c5 --steps 50 --warmup 5 > $O/${tag}_bench_rainbow_b1024_bf16.json 2>/dev/null
python scripts/prof_update.py iqn_atari_b512 bf16 > $O/${tag}_insitu_iqn_b512_bf16.txt 2>&1
python scripts/prof_update.py rainbow_atari_b1024 bf16 > $O/${tag}_insitu_rainbow_b1024_bf16.txt 2>&1
python scripts/prof_update.py dqn_atari_b32 fp32 > $O/${tag}_insitu_dqn_b32_fp32.txt 2>&1
python - <<PY
import json,glob
for f in sorted(glob.glob("$O/${tag}_bench_*.json")):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1]); print(f.split('/')[-1], d["ms_per_step"], d["value"], "e2e", d["e2e"]["value"])
    except Exception as e: print(f, "ERR", e)
PY
