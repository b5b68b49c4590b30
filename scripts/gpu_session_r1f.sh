#!/bin/bash
# Round-1f GPU session: quantisation-aware conv plans, attention staging v2; full measurement set for profiles/.
set -u
mkdir -p gpurun_out
export PYTHONDONTWRITEBYTECODE=1
nvidia-smi --query-gpu=name,driver_version,memory.total --format=csv > gpurun_out/gpu.txt 2>&1
echo "== pytest -m gpu"
timeout 900 python -m pytest tests -m gpu -q -rfE --maxfail=8 > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -n 6 gpurun_out/pytest_gpu.log | cut -c1-300
echo "== smoke"
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; echo "smoke rc=$?"; tail -n 1 gpurun_out/smoke.log
echo "== microbenchmarks"
timeout 120 python scripts/conv_bench.py > gpurun_out/convb_all.txt 2>&1; cat gpurun_out/convb_all.txt
timeout 120 python scripts/attn_bench.py > gpurun_out/attnb_new.txt 2>&1; cat gpurun_out/attnb_new.txt
timeout 120 python scripts/attn_bench.py --nk 300 > gpurun_out/attnb_new300.txt 2>&1; cat gpurun_out/attnb_new300.txt
echo "== bench"
timeout 600 python bench.py --steps 20 --warmup 3 > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "bench rc=$?"; tail -c 1200 gpurun_out/bench.json; tail -n 3 gpurun_out/bench.err
timeout 600 python bench.py --steps 20 --warmup 3 --workload ad --no-cpu-baseline > gpurun_out/bench_ad.json 2> gpurun_out/bench_ad.err; echo "bench ad rc=$?"; tail -c 600 gpurun_out/bench_ad.json
timeout 600 python bench.py --steps 3 --warmup 1 --impl reference > gpurun_out/bench_ref.json 2> gpurun_out/bench_ref.err; echo "bench ref rc=$?"; cat gpurun_out/bench_ref.json
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
du -sh gpurun_out
