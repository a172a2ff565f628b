#!/bin/bash
# A/B of k_front's work distribution (O2A_FRONT_CHUNK) on the device-resident C2 line + ncu --set full captures.
TAG=${1:-ab}; OUT=gpurun_out/$TAG; mkdir -p $OUT
B="python bench.py --no-e2e --no-cpu-baseline --no-files --no-python-reference"
for c in $2; do
  O2A_FRONT_CHUNK=$c timeout 600 $B --steps 10 --warmup 3 > $OUT/bench_chunk$c.json 2> $OUT/bench_chunk$c.err
  python - <<PY
import json
try:
    d = json.loads(open("$OUT/bench_chunk$c.json").read().strip().splitlines()[-1])
    print("chunk $c", round(d["ms_per_step"], 4), {k: round(v, 4) for k, v in d["roofline"]["stage_ms_per_step"].items()})
except Exception as e:
    print("no line", e); print(open("$OUT/bench_chunk$c.err").read()[-2000:])
PY
done
for c in $3; do
  O2A_FRONT_CHUNK=$c timeout 1200 ncu --set full --clock-control none --import-source on -k regex:'k_front' \
      --launch-skip 1 -c 1 -o $OUT/full_front_chunk$c -f $B --steps 1 --warmup 1 > $OUT/ncu_full_chunk$c.log 2>&1; echo "ncu full rc=$?"
  ncu -i $OUT/full_front_chunk$c.ncu-rep --page raw --csv > $OUT/full_front_chunk${c}_raw.csv 2>/dev/null
done
