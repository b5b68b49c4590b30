#!/bin/bash
# ncu --set full of kernels matching $1 (regex), skipping $2 launches, capturing $3; report -> gpurun_out/prof_$4.ncu-rep
set -u
mkdir -p gpurun_out
export PYTHONDONTWRITEBYTECODE=1
BENCH="python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-roofline ${BENCH_ARGS:-}"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"$1" -s $2 -c $3 -f -o gpurun_out/prof_$4 $BENCH > gpurun_out/ncu_$4.log 2>&1; echo "rc=$?"
tail -n 3 gpurun_out/ncu_$4.log
