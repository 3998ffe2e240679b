for cfg in "1 0" "0 0" "1 1" "0 1" "2 0" "3 0"; do set -- $cfg
  for prec in fp32 bf16; do
    v=$(ZOO_TC_STAGES=$1 ZOO_TMA_DGRAD=$2 python bench.py --steps 300 --warmup 5 --no-cpu --precision $prec | python -c 'import sys,json; d=json.loads(sys.stdin.read()); print("%.0f e2e %.0f"%(d["value"], d["e2e"]["value"]))')
    echo "stages_mode=$1 tma_dgrad=$2 $prec: $v"
  done
done
