#!/bin/bash
# Round-2 evidence pass: ncu --set full of the kernel families of one eager step (raw metric tables exported on the box),
# the launch list of the captured step, and the default bench line.  Everything lands in gpurun_out/${TAG}_*.
set -u
TAG=${TAG:-r2t}
mkdir -p gpurun_out
export PYTHONDONTWRITEBYTECODE=1
EAGER="python bench.py --steps 1 --warmup 3 --mode eager --no-cpu-baseline --no-roofline --no-extras"
NCU="ncu --set full --clock-control none --import-source on -f"
prof() {  # name workload regex skip count
  timeout 900 $NCU -k regex:"$3" -s $4 -c $5 -o /tmp/prof_$1 $EAGER --workload $2 > gpurun_out/${TAG}_ncu_$1.log 2>&1; echo "$1 rc=$?"
  ncu -i /tmp/prof_$1.ncu-rep --page raw --csv > gpurun_out/${TAG}_ncu_$1_raw.csv 2>/dev/null
}
prof conv cnn_ad "conv3d_" 54 18
prof block1bn cnn_ad "bn_|conv1_" ${BN_SKIP:-114} ${BN_COUNT:-38}
prof enc ad "attn_mma|enc_|token_pool" ${ENC_SKIP:-156} ${ENC_COUNT:-52}
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 8000 --csv --log-file gpurun_out/${TAG}_launches_ad.csv python bench.py --steps 2 --warmup 3 --workload ad --no-cpu-baseline --no-roofline --no-extras > gpurun_out/${TAG}_ncu_launches.log 2>&1; echo "launch list rc=$?"
timeout 900 python bench.py > gpurun_out/${TAG}_bench.json 2> gpurun_out/${TAG}_bench.err; echo "bench rc=$?"; head -c 300 gpurun_out/${TAG}_bench.json; echo
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/${TAG}_bench_ref.json 2> gpurun_out/${TAG}_bench_ref.err; echo "ref rc=$?"; head -c 300 gpurun_out/${TAG}_bench_ref.json; echo
du -sh gpurun_out
