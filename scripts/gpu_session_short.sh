#!/bin/bash
# Short GPU session: parity tests + both bench lines.
set -u
mkdir -p gpurun_out
export PYTHONDONTWRITEBYTECODE=1
timeout 600 python -m pytest tests -m gpu -q -rfE --maxfail=8 > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -n 12 gpurun_out/pytest_gpu.log | cut -c1-400
timeout 600 python bench.py --steps 20 --warmup 3 --no-cpu-baseline > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "bench rc=$?"; head -c 300 gpurun_out/bench.json; echo; tail -n 3 gpurun_out/bench.err | cut -c1-300
timeout 600 python bench.py --steps 20 --warmup 3 --workload ad --no-cpu-baseline > gpurun_out/bench_ad.json 2> gpurun_out/bench_ad.err; echo "bench ad rc=$?"; head -c 300 gpurun_out/bench_ad.json; echo
eval "${EXTRA:-true}"
