#!/bin/bash
# Round-end evidence run (one GPU): bench lines, in-situ profiles, ncu launch list and one --set full capture.
# usage: bash scripts/final_artifacts.sh <tag>   -> everything lands in gpurun_out/<tag>_*
tag=${1:-r01e}
O=gpurun_out
python bench.py > $O/${tag}_bench_dqn_b32_fp32.json 2> $O/${tag}_bench_dqn_b32_fp32.err
python bench.py --precision bf16 --no-cpu > $O/${tag}_bench_dqn_b32_bf16.json 2>/dev/null
python bench.py --impl reference --steps 3 --warmup 1 > $O/${tag}_bench_reference.json 2>/dev/null
for wl in rainbow_atari_b1024 iqn_atari_b512; do
  python bench.py --workload $wl --precision bf16 --no-cpu --steps 50 --warmup 5 > $O/${tag}_bench_${wl}_bf16.json 2>/dev/null
done
python bench.py --workload rainbow_atari_b32 --no-cpu > $O/${tag}_bench_rainbow_atari_b32_fp32.json 2>/dev/null
python bench.py --workload iqn_atari_b32 --no-cpu > $O/${tag}_bench_iqn_atari_b32_fp32.json 2>/dev/null
python bench.py --workload basic_dqn_cartpole_b32 --no-cpu > $O/${tag}_bench_basic_dqn_cartpole_fp32.json 2>/dev/null
python scripts/prof_update.py dqn_atari_b32 fp32 > $O/${tag}_insitu_dqn_b32_fp32.txt 2>&1
python scripts/prof_update.py rainbow_atari_b1024 bf16 > $O/${tag}_insitu_rainbow_b1024_bf16.txt 2>&1
python scripts/prof_update.py iqn_atari_b512 bf16 > $O/${tag}_insitu_iqn_b512_bf16.txt 2>&1
python scripts/bench_gemm.py 1,2,3 > $O/${tag}_gemm_engine_device_times.txt 2>&1
python scripts/ablate_x3.py > $O/${tag}_x3_ablation.txt 2>&1
python scripts/trace_gemm.py > $O/${tag}_gemm_phase_trace.txt 2>&1
# ncu passes last (numbers printed under ncu are never bench values)
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 420 --csv --log-file $O/${tag}_launches.csv \
    python bench.py --steps 2 --warmup 3 --no-cpu > $O/${tag}_ncu1.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:tc_gemm_kernel --launch-skip 60 -c 24 -o /tmp/${tag}_tc_full -f \
    python bench.py --steps 2 --warmup 3 --no-cpu > $O/${tag}_ncu2.log 2>&1
ncu -i /tmp/${tag}_tc_full.ncu-rep --page raw --csv > $O/${tag}_tc_raw.csv 2>/dev/null
ls -la $O | tail -30
