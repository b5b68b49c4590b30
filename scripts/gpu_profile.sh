#!/bin/bash
# ncu launch list (whole step) + full captures of the dominant kernels.  Outputs in gpurun_out/.
set -u
mkdir -p gpurun_out
export PYTHONDONTWRITEBYTECODE=1
BENCH="python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-roofline"
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 3000 --csv --log-file gpurun_out/launches.csv $BENCH > gpurun_out/ncu_launches.log 2>&1; echo "launch list rc=$?"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:conv3d_umma_kernel -s 22 -c 4 -f -o gpurun_out/prof_conv_umma $BENCH > gpurun_out/ncu_conv.log 2>&1; echo "conv rc=$?"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:conv3d_wgrad_umma_kernel -s 12 -c 2 -f -o gpurun_out/prof_wgrad_umma $BENCH > gpurun_out/ncu_wgrad.log 2>&1; echo "wgrad rc=$?"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"bn_act_pool|conv1_" -s 42 -c 6 -f -o gpurun_out/prof_membound $BENCH > gpurun_out/ncu_mem.log 2>&1; echo "membound rc=$?"
ls -la gpurun_out/*.ncu-rep
