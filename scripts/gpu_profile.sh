#!/bin/bash
# ncu --set full captures of one whole eager step's worth of each kernel family.  Outputs in gpurun_out/.
set -u
mkdir -p gpurun_out
export PYTHONDONTWRITEBYTECODE=1
BENCH="python bench.py --steps 1 --warmup 3 --mode eager --no-cpu-baseline --no-roofline ${BENCH_ARGS:-}"
NCU="ncu --set full --clock-control none --import-source on -f"
timeout 900 $NCU -k regex:conv3d_umma_kernel -s 36 -c 12 -o gpurun_out/prof_conv_umma $BENCH > gpurun_out/ncu_conv.log 2>&1; echo "conv rc=$?"
timeout 900 $NCU -k regex:conv3d_wgrad_umma_kernel -s 18 -c 6 -o gpurun_out/prof_wgrad_umma $BENCH > gpurun_out/ncu_wgrad.log 2>&1; echo "wgrad rc=$?"
timeout 900 $NCU -k regex:"bn_act_pool|conv1_" -s 69 -c 23 -o gpurun_out/prof_membound $BENCH > gpurun_out/ncu_mem.log 2>&1; echo "membound rc=$?"
ls -la gpurun_out/*.ncu-rep
