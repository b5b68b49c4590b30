#!/bin/bash
# Round-1g GPU session: in-CTA split-K Linear GEMM.
set -u
mkdir -p gpurun_out
export PYTHONDONTWRITEBYTECODE=1
timeout 600 python -m pytest tests -m gpu -q -rfE --maxfail=8 > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -n 6 gpurun_out/pytest_gpu.log | cut -c1-300
timeout 120 python scripts/linear_bench.py > gpurun_out/linb_v2.txt 2>&1; cat gpurun_out/linb_v2.txt
TMF_GEMM_IMPL=1 timeout 120 python scripts/linear_bench.py > gpurun_out/linb_v1.txt 2>&1; cat gpurun_out/linb_v1.txt
timeout 600 python bench.py --steps 20 --warmup 3 --workload ad --no-cpu-baseline > gpurun_out/bench_ad.json 2> gpurun_out/bench_ad.err; echo "bench ad rc=$?"; tail -c 400 gpurun_out/bench_ad.json
timeout 600 python bench.py --steps 20 --warmup 3 --no-cpu-baseline > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "bench rc=$?"; head -c 600 gpurun_out/bench.json
