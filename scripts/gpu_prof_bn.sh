#!/bin/bash
# ncu --set full of the BatchNorm / activation / pool passes of one eager model_CNN_ad step + launch list of a model_ad step
set -u
mkdir -p gpurun_out
export PYTHONDONTWRITEBYTECODE=1
BENCH="python bench.py --steps 1 --warmup 3 --mode eager --workload cnn_ad --no-cpu-baseline --no-roofline --no-extras"
# warm-up = 3 eager steps + 1 timed: skip the launches of the first 3 steps
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"bn_act_pool|bn_maxpool|bn_finalize|bn_bwd_finalize" -s ${SKIP:-0} -c ${COUNT:-40} -f -o gpurun_out/prof_bn $BENCH > gpurun_out/ncu_bn.log 2>&1; echo "rc=$?"
tail -n 3 gpurun_out/ncu_bn.log
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 1500 --csv --log-file gpurun_out/launches_ad_r2p.csv python bench.py --steps 1 --warmup 1 --mode eager --workload ad --no-cpu-baseline --no-roofline --no-extras > gpurun_out/ncu_launches_ad.log 2>&1; echo "rc=$?"
du -sh gpurun_out
