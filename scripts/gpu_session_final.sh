#!/bin/bash
# Full measurement set for profiles/: parity tests (abort on failure), smoke, microbenchmarks, bench lines (ours + reference
# arm), ncu launch lists, ncu --set full of the conv family and of block 1, conv1 forward bottleneck switches.
set -u
mkdir -p gpurun_out
export PYTHONDONTWRITEBYTECODE=1
nvidia-smi --query-gpu=name,driver_version,memory.total --format=csv > gpurun_out/gpu.txt 2>&1
echo "== pytest -m gpu"
timeout 900 python -m pytest tests -m gpu -q -rfE --maxfail=8 > gpurun_out/pytest_gpu.log 2>&1; rc=$?; echo "pytest rc=$rc"; tail -n 6 gpurun_out/pytest_gpu.log | cut -c1-300
if [ $rc -ne 0 ]; then echo "tests failed: stopping"; exit 1; fi
echo "== smoke"
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; echo "smoke rc=$?"; tail -n 1 gpurun_out/smoke.log
echo "== bench"
timeout 600 python bench.py --steps 20 --warmup 3 > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "bench rc=$?"; head -c 400 gpurun_out/bench.json; echo
timeout 600 python bench.py --steps 20 --warmup 3 --workload ad --no-cpu-baseline > gpurun_out/bench_ad.json 2> gpurun_out/bench_ad.err; echo "bench ad rc=$?"; head -c 300 gpurun_out/bench_ad.json; echo
timeout 600 python bench.py --steps 3 --warmup 1 --impl reference > gpurun_out/bench_ref.json 2> gpurun_out/bench_ref.err; echo "bench ref rc=$?"; head -c 300 gpurun_out/bench_ref.json; echo
echo "== microbenchmarks"
timeout 120 python scripts/conv_bench.py > gpurun_out/convb_all.txt 2>&1; cat gpurun_out/convb_all.txt
timeout 120 python scripts/attn_bench.py > gpurun_out/attnb_new.txt 2>&1; cat gpurun_out/attnb_new.txt
timeout 120 python scripts/linear_bench.py > gpurun_out/linb_v2.txt 2>&1; tail -n 1 gpurun_out/linb_v2.txt
echo "== ncu launch lists (eager, one step)"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 1500 --csv --log-file gpurun_out/launches.csv \
  python bench.py --steps 1 --warmup 3 --mode eager --no-cpu-baseline --no-roofline > gpurun_out/ncu_bench.log 2>&1; echo "ncu rc=$?"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 3000 --csv --log-file gpurun_out/launches_ad.csv \
  python bench.py --steps 1 --warmup 3 --mode eager --workload ad --no-cpu-baseline --no-roofline > gpurun_out/ncu_bench_ad.log 2>&1; echo "ncu ad rc=$?"
echo "== ncu --set full"
NCU="ncu --set full --clock-control none --import-source on -f"
BENCH="python bench.py --steps 1 --warmup 2 --mode eager --no-cpu-baseline --no-roofline"
prof() {
  timeout 600 $NCU -k regex:"$2" -s $3 -c $4 -o gpurun_out/prof_$1 $BENCH > gpurun_out/ncu_$1.log 2>&1; echo "$1 rc=$?"
  ncu -i gpurun_out/prof_$1.ncu-rep --page raw --csv > gpurun_out/ncu_$1_raw.csv 2>/dev/null
}
prof conv "conv3d_umma" 24 12
prof wgrad "conv3d_wgrad" 12 6
prof block1 "conv1_umma_fwd|bn_act_pool_fwd|reduce_kept|conv1_bwd_fused" 24 12
rm -f gpurun_out/prof_block1.ncu-rep gpurun_out/prof_wgrad.ncu-rep      # keep the merge under 64 MiB: raw CSVs stay
echo "== conv1 forward timing switches"
for d in 0 1 2 4 7; do TMF_C1U_DEBUG=$d timeout 60 python scripts/conv1_bench.py >> gpurun_out/conv1b.txt 2>&1; done; cat gpurun_out/conv1b.txt
du -sh gpurun_out
