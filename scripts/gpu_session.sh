#!/bin/bash
# GPU-box session: parity tests, smoke, bench lines, microbenchmarks.  Everything lands in gpurun_out/<TAG>_*.
#   TAG=r2a RUN_TESTS=1 RUN_BENCH=1 RUN_MICRO=1 bash scripts/gpu_session.sh
set -u
TAG=${TAG:-r2}
mkdir -p gpurun_out
export PYTHONDONTWRITEBYTECODE=1
nvidia-smi --query-gpu=name,driver_version,memory.total --format=csv > gpurun_out/${TAG}_gpu.txt 2>&1
nproc >> gpurun_out/${TAG}_gpu.txt
if [ "${RUN_TESTS:-1}" = "1" ]; then
echo "== pytest -m gpu ${PYTEST_ARGS:-}"
timeout ${TEST_TIMEOUT:-1500} python -m pytest tests -m gpu -q -s -rfE --maxfail=${MAXFAIL:-30} ${PYTEST_ARGS:-} > gpurun_out/${TAG}_pytest_gpu.log 2>&1; echo "pytest rc=$?"
grep -E "^\[parity|passed|failed|^FAILED|^ERROR|\[dp\]" gpurun_out/${TAG}_pytest_gpu.log | cut -c1-1500 | tail -n 60
echo "== smoke"
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/${TAG}_smoke.log 2>&1; echo "smoke rc=$?"; tail -n 2 gpurun_out/${TAG}_smoke.log
fi
if [ "${RUN_BENCH:-1}" = "1" ]; then
echo "== bench"
timeout 900 python bench.py --steps ${BENCH_STEPS:-20} --warmup 3 ${BENCH_ARGS:-} > gpurun_out/${TAG}_bench.json 2> gpurun_out/${TAG}_bench.err; echo "bench rc=$?"; head -c 600 gpurun_out/${TAG}_bench.json; echo; tail -n 3 gpurun_out/${TAG}_bench.err
fi
if [ "${RUN_MICRO:-0}" = "1" ]; then
echo "== microbenchmarks"
timeout 200 python scripts/conv_bench.py > gpurun_out/${TAG}_convb.txt 2>&1; cat gpurun_out/${TAG}_convb.txt
fi
if [ -n "${EXTRA:-}" ]; then
echo "== extra: $EXTRA"
bash -c "$EXTRA"
fi
du -sh gpurun_out
