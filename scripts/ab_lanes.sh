run() { env "$@" timeout 200 python bench.py --no-cpu --no-c5 --steps 200 --warmup 20 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('$*', round(d['value'],1), round(d['ms_per_step'],4), round(d['e2e']['value'],1))"; }
run A=1
run ZOO_MIDK_SPLIT=2
run ZOO_MIDK_SPLIT=3
run ZOO_DQN_TAIL=2
