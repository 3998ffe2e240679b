#!/bin/bash
# A/B of one environment knob on the same box: scripts/ab_env.sh NAME v1 v2 ... (fp32 default bench, 2 repeats)
name=$1; shift
for rep in 1 2; do
  for v in "$@"; do
    r=$(env $name=$v python bench.py --steps 300 --warmup 5 --no-cpu | python -c 'import sys,json; d=json.loads(sys.stdin.read()); print("%.0f e2e %.0f"%(d["value"], d["e2e"]["value"]))')
    echo "$name=$v: $r"
  done
done
