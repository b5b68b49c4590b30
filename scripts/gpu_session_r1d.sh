#!/bin/bash
# Round-1d GPU session: new kernels (multi-lane TMA + two-issuer conv, pipelined GEMM): parity, microbenchmarks,
# both bench lines, model_ad launch list, --set full of the block-1 (HBM-bound) kernels.
set -u
mkdir -p gpurun_out
export PYTHONDONTWRITEBYTECODE=1
echo "== pytest -m gpu"
timeout 900 python -m pytest tests -m gpu -q -x -rfE > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -n 4 gpurun_out/pytest_gpu.log
echo "== conv microbench (layers 4,5)"
timeout 120 python scripts/conv_bench.py --layers 4,5 --ops fwd,dgrad > gpurun_out/convb_new.txt 2>&1; cat gpurun_out/convb_new.txt
TMF_UMMA_ISSUERS=1 timeout 120 python scripts/conv_bench.py --layers 4,5 --ops fwd,dgrad > gpurun_out/convb_iss1.txt 2>&1; cat gpurun_out/convb_iss1.txt
TMF_UMMA_ISSUERS=1 TMF_UMMA_TMA_LANES=1 timeout 120 python scripts/conv_bench.py --layers 4,5 --ops fwd,dgrad > gpurun_out/convb_old.txt 2>&1; cat gpurun_out/convb_old.txt
echo "== linear microbench"
timeout 120 python scripts/linear_bench.py > gpurun_out/linb_new.txt 2>&1; cat gpurun_out/linb_new.txt
TMF_GEMM_IMPL=0 timeout 120 python scripts/linear_bench.py > gpurun_out/linb_old.txt 2>&1; cat gpurun_out/linb_old.txt
echo "== bench"
timeout 600 python bench.py --steps 10 --warmup 3 > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "bench rc=$?"; tail -c 1500 gpurun_out/bench.json
timeout 600 python bench.py --steps 10 --warmup 3 --workload ad --no-cpu-baseline > gpurun_out/bench_ad.json 2> gpurun_out/bench_ad.err; echo "bench ad rc=$?"; tail -c 600 gpurun_out/bench_ad.json
echo "== ncu launch list, model_ad eager"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 3000 --csv --log-file gpurun_out/launches_ad.csv \
  python bench.py --steps 1 --warmup 3 --mode eager --workload ad --no-cpu-baseline --no-roofline > gpurun_out/ncu_bench_ad.log 2>&1; echo "ncu rc=$?"
echo "== ncu --set full, block 1"
BENCH="python bench.py --steps 1 --warmup 2 --mode eager --no-cpu-baseline --no-roofline"
timeout 600 ncu --set full --clock-control none --import-source on -f -k regex:"conv1_umma_fwd|bn_act_pool|conv1_bwd_fused" -s 44 -c 22 \
  -o gpurun_out/prof_block1 $BENCH > gpurun_out/ncu_block1.log 2>&1; echo "block1 rc=$?"
ncu -i gpurun_out/prof_block1.ncu-rep --page raw --csv > gpurun_out/ncu_block1_raw.csv 2>/dev/null
du -sh gpurun_out
