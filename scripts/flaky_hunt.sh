#!/bin/bash
# repeats one parity test under an environment setting and counts failures: scripts/flaky_hunt.sh N "ENV=.. ENV2=.." <pytest -k expr>
n=$1; envs=$2; k=$3; fail=0
for i in $(seq 1 $n); do
  env $envs python -m pytest tests/test_gpu_learners.py -m gpu -q -x -k "$k" > /tmp/fh.log 2>&1 || { fail=$((fail+1)); grep -E "rel err|Error" /tmp/fh.log | head -2; }
done
echo "[$envs] $k: $fail / $n failed"
