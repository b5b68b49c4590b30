#!/bin/bash
# bench in both modes (no CPU baseline)
set -u
mkdir -p gpurun_out
export PYTHONDONTWRITEBYTECODE=1
for mode in graph eager; do
timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --mode $mode ${BENCH_ARGS:-} > gpurun_out/bench_$mode.json 2> gpurun_out/bench_$mode.err; echo "bench $mode rc=$?"
python - $mode <<'PY'
import json, sys
m = sys.argv[1]
try:
    d = json.loads(open(f"gpurun_out/bench_{m}.json").read().strip().splitlines()[-1])
    print(m, "value", round(d["value"], 1), "ms", round(d["ms_per_step"], 3), "e2e", round(d["e2e"]["value"], 1), "e2e ms", round(d["e2e"]["ms_per_step"], 3), "launches", d["gpu_launches"], "roofline", d["roofline"] and round(d["roofline"]["frac"], 4))
except Exception as e:
    print("no bench line:", e)
    print(open(f"gpurun_out/bench_{m}.err").read()[-3000:])
PY
done
