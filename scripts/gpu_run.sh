#!/bin/bash
# Standard GPU-box session: parity tests, smoke, bench, launch list.  Everything lands in gpurun_out/.
set -u
mkdir -p gpurun_out
export PYTHONDONTWRITEBYTECODE=1
nvidia-smi --query-gpu=name,driver_version,memory.total --format=csv > gpurun_out/gpu.txt 2>&1
echo "== pytest -m gpu" 
timeout 1200 python -m pytest tests -m gpu -q -s -rfE --maxfail=40 > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"
tail -n 60 gpurun_out/pytest_gpu.log
echo "== smoke"
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; echo "smoke rc=$?"; tail -n 5 gpurun_out/smoke.log
echo "== bench"
timeout 900 python bench.py --steps ${BENCH_STEPS:-5} --warmup 3 > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "bench rc=$?"; tail -c 3000 gpurun_out/bench.json; tail -n 5 gpurun_out/bench.err
if [ "${RUN_AD:-1}" = "1" ]; then
timeout 900 python bench.py --steps ${BENCH_STEPS:-5} --warmup 3 --workload ad --no-cpu-baseline > gpurun_out/bench_ad.json 2> gpurun_out/bench_ad.err; echo "bench ad rc=$?"; tail -c 2500 gpurun_out/bench_ad.json; tail -n 5 gpurun_out/bench_ad.err
fi
if [ "${RUN_NCU:-1}" = "1" ]; then
echo "== ncu launch list"
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 1500 --csv --log-file gpurun_out/launches.csv \
  python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-roofline > gpurun_out/ncu_bench.log 2>&1; echo "ncu rc=$?"
fi
