#!/bin/bash
# Round-1c GPU session: parity tests, smoke, both bench lines, eager launch list, one --set full capture per conv family.
set -u
mkdir -p gpurun_out
export PYTHONDONTWRITEBYTECODE=1
BENCH_STEPS=10 RUN_NCU=0 bash scripts/gpu_run.sh
echo "== ncu launch list (eager, one step)"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 1500 --csv --log-file gpurun_out/launches.csv \
  python bench.py --steps 1 --warmup 3 --mode eager --no-cpu-baseline --no-roofline > gpurun_out/ncu_bench.log 2>&1; echo "ncu rc=$?"
echo "== ncu --set full"
NCU="ncu --set full --clock-control none --import-source on -f"
BENCH="python bench.py --steps 1 --warmup 2 --mode eager --no-cpu-baseline --no-roofline"
prof() {
  timeout 600 $NCU -k regex:"$2" -s $3 -c $4 -o gpurun_out/prof_$1 $BENCH > gpurun_out/ncu_$1.log 2>&1; echo "$1 rc=$?"
  ncu -i gpurun_out/prof_$1.ncu-rep --page raw --csv > gpurun_out/ncu_$1_raw.csv 2>/dev/null
}
# skip the two warm-up steps' launches: capture the timed step's 12 fwd/dgrad launches and 6 wgrad launches
prof conv "conv3d_umma" 24 12
prof wgrad "conv3d_wgrad" 12 6
du -sh gpurun_out
