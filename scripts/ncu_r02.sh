#!/bin/bash
# Round-2 ncu evidence (ONE gpurun call, one GPU).  Every capture is preceded by the same command without ncu.
# usage: bash scripts/ncu_r02.sh <tag>  -> gpurun_out/<tag>_ncu_*.csv (raw pages exported on the box) + launch list
tag=${1:-r02}
O=gpurun_out
full() {  # <name> <workload> <precision>
  python scripts/one_update.py $2 $3 1 3 > $O/${tag}_plain_$1.log 2>&1 &&
  timeout 900 ncu --set full --clock-control none --import-source on --profile-from-start off -o /tmp/${tag}_$1 -f \
      python scripts/one_update.py $2 $3 1 3 > $O/${tag}_ncu_$1.log 2>&1
  ncu -i /tmp/${tag}_$1.ncu-rep --page raw --csv > $O/${tag}_ncu_$1.csv 2>/dev/null
}
full dqn_b32_fp32 dqn_atari_b32 fp32
full rainbow_b1024_bf16 rainbow_atari_b1024 bf16
full iqn_b512_bf16 iqn_atari_b512 bf16
python scripts/probe_sumtree_small.py > $O/${tag}_plain_sumtree.log 2>&1 &&
timeout 300 ncu --set full --clock-control none -k regex:sumtree -c 12 -o /tmp/${tag}_sumtree -f python scripts/probe_sumtree_small.py > $O/${tag}_ncu_sumtree.log 2>&1
ncu -i /tmp/${tag}_sumtree.ncu-rep --page raw --csv > $O/${tag}_ncu_sumtree.csv 2>/dev/null
python bench.py --steps 2 --warmup 3 --no-cpu --no-c5 > $O/${tag}_plain_bench.log 2>&1 &&
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 420 --csv --log-file $O/${tag}_launches.csv \
    python bench.py --steps 2 --warmup 3 --no-cpu --no-c5 > $O/${tag}_ncu_launches.log 2>&1
ls -la $O | grep ${tag}_ | tail -20
