#!/bin/bash
# The BASELINE configs at stated size + KDE + the fixed-workload (strong) mode on ONE GPU. Usage: bash tools/gpu_configs.sh <tag>
TAG=${1:-cfg}; OUT=gpurun_out/$TAG; mkdir -p $OUT
run() { name=$1; shift; t0=$(date +%s); timeout 1200 python bench.py "$@" > $OUT/$name.json 2> $OUT/$name.err; echo "$name rc=$? $(( $(date +%s) - t0 )) s; $(tail -1 $OUT/$name.err | head -c 300)"; head -c 500 $OUT/$name.json; echo; }
run c2_hist --steps 10 --warmup 3
run c2_kde --fitter kde --steps 10 --warmup 3 --cpu-lnc 300
run c5_hist --config 5 --steps 10 --warmup 3
run c5_strong_n1 --scaling strong --steps 20 --warmup 3
run c1_hist --config 1 --steps 20 --warmup 3
run c4_full --config 4 --steps 3 --warmup 1 --cpu-lnc 8 --e2e-steps 2
run c3_hist_full --config 3 --lnc 1000 --batches 20 --steps 5 --warmup 2 --cpu-lnc 200
run c3_kde --config 3 --fitter kde --lnc 1000 --batches 4 --steps 5 --warmup 2 --cpu-lnc 2
