#!/bin/bash
# Quick GPU check: op + model parity tests, then one bench line (no CPU baseline).  Outputs in gpurun_out/.
set -u
mkdir -p gpurun_out
export PYTHONDONTWRITEBYTECODE=1
timeout 900 python -m pytest tests -m gpu -q -x -rfE ${PYTEST_ARGS:-} > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"
tail -n ${TAILN:-15} gpurun_out/pytest_gpu.log
timeout 600 python bench.py --steps 5 --warmup 3 --no-cpu-baseline ${BENCH_ARGS:-} > gpurun_out/bench_quick.json 2> gpurun_out/bench_quick.err; echo "bench rc=$?"
python - <<'PY'
import json
try:
    d = json.loads(open("gpurun_out/bench_quick.json").read().strip().splitlines()[-1])
    print("value", round(d["value"], 1), "ms", round(d["ms_per_step"], 3), "e2e", round(d["e2e"]["value"], 1), "roofline", d["roofline"] and round(d["roofline"]["frac"], 4))
    b = d.get("breakdown") or {}
    print(" ".join(f"{k.replace('tmf_','')}={v}" for k, v in b.items()))
except Exception as e:
    print("no bench line:", e)
    print(open("gpurun_out/bench_quick.err").read()[-2000:])
PY
