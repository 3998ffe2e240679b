#!/bin/bash
# compute-sanitizer pass over the engine, SumTree and learner parity tests: scripts/sanitize.sh <memcheck|racecheck|synccheck|initcheck> [pytest -k expr]
# One tool per gpurun call (B200_PROFILING.md).  Writes gpurun_out/r02_sanitizer_<tool>.{log,txt}.
tool=${1:-memcheck}; kexpr=${2:-"not non_default and not bf16 and not simt"}
out=gpurun_out/r02_sanitizer_$tool
timeout ${SAN_TIMEOUT:-1200} compute-sanitizer --tool $tool --print-limit 40 --log-file $out.log \
  python -m pytest tests/test_gpu_gemm.py tests/test_gpu_sumtree.py tests/test_gpu_learners.py tests/test_replay.py -m gpu -q -x -k "$kexpr" > $out.txt 2>&1
echo "rc=$?" >> $out.txt
tail -5 $out.txt
grep -E "ERROR SUMMARY|RACECHECK SUMMARY|========= (Invalid|Race|Barrier|Uninit)" $out.log | sort | uniq -c | head -40
