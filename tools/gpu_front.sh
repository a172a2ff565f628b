#!/bin/bash
# Front-kernel iteration: parity tests, then the device-resident C2 line for hist and kde, then one --set full capture.
# Usage (under gpurun): bash tools/gpu_front.sh <tag> [tests|notests] [ncu|noncu]
TAG=${1:-f}; TESTS=${2:-tests}; NCU=${3:-ncu}
OUT=gpurun_out/$TAG; mkdir -p $OUT
if [ "$TESTS" = "tests" ]; then
  timeout 1500 python -m pytest tests -m gpu -x -q > $OUT/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> $OUT/pytest_gpu.log
  tail -15 $OUT/pytest_gpu.log
fi
for fit in hist kde; do
  timeout 600 python bench.py --steps 10 --warmup 3 --fitter $fit --no-e2e --no-cpu-baseline --no-files --no-python-reference > $OUT/bench_$fit.json 2> $OUT/bench_$fit.err; echo "bench $fit rc=$?"
  python - <<PY
import json
try:
    d = json.loads(open("$OUT/bench_$fit.json").read().strip().splitlines()[-1])
    print("$fit", d["ms_per_step"], d["roofline"]["frac"], d["roofline"]["stage_ms_per_step"])
except Exception as e:
    print("no line", e); print(open("$OUT/bench_$fit.err").read()[-2000:])
PY
done
if [ "$NCU" = "ncu" ]; then
  timeout 1200 ncu --set full --clock-control none --import-source on -k regex:'k_front' \
      --launch-skip 3 -c 1 -o $OUT/full_front -f python bench.py --steps 1 --warmup 1 --no-e2e --no-cpu-baseline --no-files --no-python-reference > $OUT/ncu_full.log 2>&1; echo "ncu full rc=$?"
  ncu -i $OUT/full_front.ncu-rep --page raw --csv > $OUT/full_front_raw.csv 2>/dev/null
fi
