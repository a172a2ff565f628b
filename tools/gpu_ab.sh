#!/bin/bash
# Quick GPU round: parity tests, the device-resident arm (A/B via env), ncu --set full of the top kernels
TAG=${1:-ab}; OUT=gpurun_out/$TAG; mkdir -p $OUT
timeout 1500 python -m pytest tests -m gpu -x -q > $OUT/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -2 $OUT/pytest_gpu.log
for V in "O2A_FRONT_PREFETCH=1" "O2A_FRONT_PREFETCH=0"; do
env $V timeout 600 python bench.py --steps 10 --warmup 3 --no-e2e --no-cpu-baseline > $OUT/bench_$V.json 2> $OUT/bench_$V.err; echo "$V rc=$?"
python -c "
import json,sys
d=json.load(open('$OUT/bench_$V.json')); print('$V', d['ms_per_step'], d['roofline']['stage_ms_per_step'], d['roofline']['frac'])"
done
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'k_front|k_chain_warp|k_chain_small' --launch-skip 6 -c 3 -o $OUT/full -f python bench.py --steps 2 --warmup 1 --no-e2e --no-cpu-baseline > $OUT/ncu_full.log 2>&1; echo "ncu rc=$?"
