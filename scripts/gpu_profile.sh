#!/bin/bash
# ncu --set full captures of one whole eager step's worth of each kernel family; the raw metric tables are exported on the
# box (gpurun_out/ncu_*.csv), the bulky .ncu-rep files are kept only while gpurun_out stays small.
set -u
mkdir -p gpurun_out
export PYTHONDONTWRITEBYTECODE=1
BENCH="python bench.py --steps 1 --warmup 3 --mode eager --no-cpu-baseline --no-roofline ${BENCH_ARGS:-}"
NCU="ncu --set full --clock-control none --import-source on -f"
prof() {  # name regex skip count
  timeout 900 $NCU -k regex:"$2" -s $3 -c $4 -o /tmp/prof_$1 $BENCH > gpurun_out/ncu_$1.log 2>&1; echo "$1 rc=$?"
  ncu -i /tmp/prof_$1.ncu-rep --page raw --csv > gpurun_out/ncu_$1_raw.csv 2>/dev/null
  ncu -i /tmp/prof_$1.ncu-rep --page details --csv > gpurun_out/ncu_$1_details.csv 2>/dev/null
}
prof conv "conv3d_umma" ${CONV_SKIP:-36} ${CONV_COUNT:-12}
prof wgrad "conv3d_wgrad_umma_kernel" 18 6
prof mem "bn_act_pool|conv1_" 69 23
if [ -n "${KEEP_REP:-}" ]; then cp /tmp/prof_${KEEP_REP}.ncu-rep gpurun_out/ 2>/dev/null; fi
du -sh gpurun_out
