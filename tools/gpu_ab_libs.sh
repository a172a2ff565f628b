#!/bin/bash
# A/B of library builds (ab_libs/<name>.so, compile-time variants) x environment settings on the device-resident line.
# Usage (under gpurun): bash tools/gpu_ab_libs.sh <tag> "<bench args>" "<lib>:<ENV=V,ENV=V>" ...   (lib "-" = the in-tree build)
TAG=$1; ARGS=$2; shift 2; OUT=gpurun_out/$TAG; mkdir -p $OUT
cp ortho2align_b200/lib/libo2a_cuda.so $OUT/.orig.so
B="python bench.py --no-e2e --no-cpu-baseline --no-files --no-python-reference --no-strong $ARGS"
for spec in "$@"; do
  lib=${spec%%:*}; envs=${spec#*:}; [ "$envs" = "$spec" ] && envs=""
  if [ "$lib" = "-" ]; then cp $OUT/.orig.so ortho2align_b200/lib/libo2a_cuda.so; else cp ab_libs/$lib.so ortho2align_b200/lib/libo2a_cuda.so; fi
  name=$(echo "${lib}_${envs}" | tr ',=' '__')
  env $(echo $envs | tr ',' ' ') timeout 600 $B > $OUT/$name.json 2> $OUT/$name.err
  python - <<PY
import json
try:
    d = json.loads(open("$OUT/$name.json").read().strip().splitlines()[-1])
    print("$spec", round(d["ms_per_step"], 4), {k: round(v, 4) for k, v in d["roofline"]["stage_ms_per_step"].items()})
except Exception as e:
    print("$spec: no line", e); print(open("$OUT/$name.err").read()[-1500:])
PY
done
cp $OUT/.orig.so ortho2align_b200/lib/libo2a_cuda.so; rm -f $OUT/.orig.so
