#!/bin/bash
# A/B of two builds of libzoo_sm100 on the same box: scripts/ab_libs.sh <other.so> [bench args]
other=$1; shift
for rep in 1 2; do
  for lib in "" "$other"; do
    for prec in fp32 bf16; do
      v=$(ZOO_SM100_LIB=$lib python bench.py --steps 300 --warmup 5 --no-cpu --precision $prec "$@" | python -c 'import sys,json; d=json.loads(sys.stdin.read()); print("%.0f e2e %.0f"%(d["value"], d["e2e"]["value"]))')
      echo "lib=${lib:-current} $prec: $v"
    done
  done
done
