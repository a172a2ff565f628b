#!/bin/bash
# A/B of library environment switches on the device-resident bench line.
# Usage (under gpurun): bash tools/gpu_env_ab.sh <tag> "<bench args>" "ENV1=a ENV2=b" "ENV1=c" ...
TAG=$1; ARGS=$2; shift 2
OUT=gpurun_out/$TAG; mkdir -p $OUT
B="python bench.py --no-e2e --no-cpu-baseline --no-files --no-python-reference --steps 10 --warmup 3 $ARGS"
i=0
for combo in "$@"; do
  i=$((i+1))
  env $combo timeout 600 $B > $OUT/bench_$i.json 2> $OUT/bench_$i.err
  python - <<PY
import json
try:
    d = json.loads(open("$OUT/bench_$i.json").read().strip().splitlines()[-1])
    print("$combo |", round(d["ms_per_step"], 4), {k: round(v, 4) for k, v in d["roofline"]["stage_ms_per_step"].items()})
except Exception as e:
    print("$combo | no line", e); print(open("$OUT/bench_$i.err").read()[-1500:])
PY
done
