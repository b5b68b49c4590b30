#!/bin/bash
# Round-1e GPU session: M-blocked streamed-weight conv plans, register-tiled attention, lagged loss reads.
set -u
mkdir -p gpurun_out
export PYTHONDONTWRITEBYTECODE=1
echo "== pytest -m gpu"
timeout 900 python -m pytest tests -m gpu -q -rfE --maxfail=8 > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -n 12 gpurun_out/pytest_gpu.log | cut -c1-400
echo "== conv microbench (layers 4,5)"
timeout 120 python scripts/conv_bench.py --layers 4,5 --ops fwd,dgrad > gpurun_out/convb_new.txt 2>&1; cat gpurun_out/convb_new.txt
TMF_UMMA_ISSUERS=1 timeout 120 python scripts/conv_bench.py --layers 4,5 --ops dgrad > gpurun_out/convb_iss1.txt 2>&1; cat gpurun_out/convb_iss1.txt
TMF_UMMA_MT=2 timeout 120 python scripts/conv_bench.py --layers 4,5 --ops fwd,dgrad > gpurun_out/convb_mt2.txt 2>&1; cat gpurun_out/convb_mt2.txt
TMF_UMMA_MT=1 timeout 120 python scripts/conv_bench.py --layers 4,5 --ops dgrad > gpurun_out/convb_mt1.txt 2>&1; cat gpurun_out/convb_mt1.txt
echo "== attention microbench"
timeout 120 python scripts/attn_bench.py > gpurun_out/attnb_new.txt 2>&1; cat gpurun_out/attnb_new.txt
TMF_ATTN_IMPL=0 timeout 120 python scripts/attn_bench.py > gpurun_out/attnb_old.txt 2>&1; cat gpurun_out/attnb_old.txt
echo "== bench"
timeout 600 python bench.py --steps 10 --warmup 3 > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "bench rc=$?"; tail -c 1200 gpurun_out/bench.json; tail -n 3 gpurun_out/bench.err
timeout 600 python bench.py --steps 10 --warmup 3 --workload ad --no-cpu-baseline > gpurun_out/bench_ad.json 2> gpurun_out/bench_ad.err; echo "bench ad rc=$?"; tail -c 600 gpurun_out/bench_ad.json; tail -n 3 gpurun_out/bench_ad.err
