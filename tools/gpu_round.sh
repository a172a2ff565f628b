#!/bin/bash
# One GPU-box round: parity tests, the default bench line, an ncu launch list and --set full captures of the top kernels.
# Usage (under gpurun): bash tools/gpu_round.sh <tag> [tests|notests] [ncu|noncu]
TAG=${1:-r2}; TESTS=${2:-tests}; NCU=${3:-ncu}
OUT=gpurun_out/$TAG; mkdir -p $OUT
nvidia-smi --query-gpu=name,clocks.max.sm,clocks.sm,memory.total --format=csv > $OUT/gpu.txt 2>&1
nproc >> $OUT/gpu.txt
if [ "$TESTS" = "tests" ]; then
  timeout 1500 python -m pytest tests -m gpu -x -q > $OUT/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> $OUT/pytest_gpu.log
  tail -5 $OUT/pytest_gpu.log
fi
timeout 900 python bench.py --steps 10 --warmup 3 > $OUT/bench_c2.json 2> $OUT/bench_c2.err; echo "bench rc=$?"
tail -c 3000 $OUT/bench_c2.json
if [ "$NCU" = "ncu" ]; then
  timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $OUT/launches_c2.csv \
      python bench.py --steps 2 --warmup 1 --no-e2e --no-cpu-baseline > $OUT/ncu_launch.log 2>&1; echo "ncu list rc=$?"
  timeout 1200 ncu --set full --clock-control none --import-source on -k regex:'k_front|k_chain_warp|k_fit_count|k_fit_hist_tail|k_gather_retained' \
      --launch-skip 10 -c 5 -o $OUT/full_c2 -f python bench.py --steps 2 --warmup 1 --no-e2e --no-cpu-baseline > $OUT/ncu_full.log 2>&1; echo "ncu full rc=$?"
  ncu -i $OUT/full_c2.ncu-rep --page raw --csv > $OUT/full_c2_raw.csv 2>/dev/null
  # the KDE branch: k_fit<int> (value compression + brentq) and k_front<double, 1> (class table by value-compressed ndtr sums)
  timeout 1200 ncu --set full --clock-control none --import-source on -k regex:'k_front|k_fit' \
      --launch-skip 6 -c 4 -o $OUT/full_c2_kde -f python bench.py --fitter kde --steps 2 --warmup 1 --no-e2e --no-cpu-baseline > $OUT/ncu_full_kde.log 2>&1; echo "ncu full kde rc=$?"
  ncu -i $OUT/full_c2_kde.ncu-rep --page raw --csv > $OUT/full_c2_kde_raw.csv 2>/dev/null
fi
