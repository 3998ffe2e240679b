#!/bin/bash
# Shorter round-end evidence run: default bench (both precisions + reference arm), in-situ profile, ncu launch list + one --set full pass.
tag=${1:-r01f}
O=gpurun_out
python bench.py > $O/${tag}_bench_dqn_b32_fp32.json 2> $O/${tag}_bench_dqn_b32_fp32.err
python bench.py --precision bf16 --no-cpu > $O/${tag}_bench_dqn_b32_bf16.json 2>/dev/null
python bench.py --impl reference --steps 3 --warmup 1 > $O/${tag}_bench_reference.json 2>/dev/null
python scripts/prof_update.py dqn_atari_b32 fp32 > $O/${tag}_insitu_dqn_b32_fp32.txt 2>&1
python scripts/bench_gemm.py 1,2,3 > $O/${tag}_gemm_engine_device_times.txt 2>&1
python scripts/trace_gemm.py > $O/${tag}_gemm_phase_trace.txt 2>&1
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 420 --csv --log-file $O/${tag}_launches.csv \
    python bench.py --steps 2 --warmup 3 --no-cpu > $O/${tag}_ncu1.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:tc_gemm_kernel --launch-skip 60 -c 24 -o /tmp/${tag}_tc_full -f \
    python bench.py --steps 2 --warmup 3 --no-cpu > $O/${tag}_ncu2.log 2>&1
ncu -i /tmp/${tag}_tc_full.ncu-rep --page raw --csv > $O/${tag}_tc_raw.csv 2>/dev/null
ls -la $O | grep ${tag}
