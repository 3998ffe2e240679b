run() { env "$@" timeout 200 python bench.py --no-cpu --steps 200 --warmup 20 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('$*', round(d['value'],1), round(d['ms_per_step'],4), round(d['e2e']['value'],1))"; }
run A=1
run ZOO_TC_STAGES=2
run ZOO_SPLIT_CAP_PCT=18
run ZOO_WGRAD_CAP_PCT=35
run ZOO_LANE_PRIO=0
