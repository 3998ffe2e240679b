#!/bin/bash
# Round-2 evidence run of the final build (one GPU): every bench line, in-situ per-launch profiles, the graph-replay phase trace,
# SumTree probe, full-size parity table.   usage: bash scripts/final_r02.sh <tag>  -> gpurun_out/<tag>_*
tag=${1:-r02}
O=gpurun_out
python bench.py > $O/${tag}_bench_dqn_b32_fp32.json 2> $O/${tag}_bench_dqn_b32_fp32.err
python bench.py --impl reference --steps 3 --warmup 1 > $O/${tag}_bench_reference.json 2>/dev/null
python bench.py --precision bf16 --no-cpu --no-c5 > $O/${tag}_bench_dqn_b32_bf16.json 2>/dev/null
for wl in rainbow_atari_b1024 iqn_atari_b512; do
  python bench.py --workload $wl --precision bf16 --no-cpu --no-c5 --steps 50 --warmup 5 > $O/${tag}_bench_${wl}_bf16.json 2>/dev/null
done
python bench.py --workload iqn_atari_b512 --no-cpu --no-c5 --steps 20 --warmup 3 > $O/${tag}_bench_iqn_atari_b512_fp32.json 2>/dev/null
python bench.py --workload rainbow_atari_b1024 --no-cpu --no-c5 --steps 30 --warmup 3 > $O/${tag}_bench_rainbow_atari_b1024_fp32.json 2>/dev/null
for wl in rainbow_atari_b32 iqn_atari_b32 basic_dqn_cartpole_b32; do
  python bench.py --workload $wl --no-cpu --no-c5 > $O/${tag}_bench_${wl}_fp32.json 2>/dev/null
done
python scripts/prof_update.py dqn_atari_b32 fp32 > $O/${tag}_insitu_dqn_b32_fp32.txt 2>&1
python scripts/prof_update.py rainbow_atari_b1024 bf16 > $O/${tag}_insitu_rainbow_b1024_bf16.txt 2>&1
python scripts/prof_update.py iqn_atari_b512 bf16 > $O/${tag}_insitu_iqn_b512_bf16.txt 2>&1
python scripts/prof_update.py iqn_atari_b512 fp32 > $O/${tag}_insitu_iqn_b512_fp32.txt 2>&1
python scripts/trace_update.py dqn_atari_b32 fp32 > $O/${tag}_phase_trace_dqn_b32_fp32.txt 2>&1
python scripts/probe_sumtree_small.py > $O/${tag}_sumtree_small.txt 2>&1
python scripts/full_size_parity.py > $O/${tag}_full_size_parity.txt 2>&1
ls -la $O | grep ${tag}_
