#!/bin/bash
# Quick GPU round: parity tests, the device-resident arm, ncu --set full of the top kernels
TAG=${1:-ab}; OUT=gpurun_out/$TAG; mkdir -p $OUT
timeout 1500 python -m pytest tests -m gpu -x -q > $OUT/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -4 $OUT/pytest_gpu.log
timeout 600 python bench.py --steps 10 --warmup 3 --no-e2e --no-cpu-baseline > $OUT/bench.json 2> $OUT/bench.err; echo "rc=$?"
python -c "
import json,sys
d=json.load(open('$OUT/bench.json')); print('bench', d['ms_per_step'], d['roofline']['stage_ms_per_step'], d['roofline']['frac'])"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'k_front|k_chain_warp|k_chain_small|k_gather_retained|k_fit_count|k_fit_hist_tail' --launch-skip 8 -c 8 -o $OUT/full -f python bench.py --steps 2 --warmup 1 --no-e2e --no-cpu-baseline > $OUT/ncu_full.log 2>&1; echo "ncu rc=$?"
