# A/B of the e2e arm under an environment knob (scratch tool): bash tools/ab_e2e.sh VAR v1 v2 ...
VAR=$1; shift
for v in "$@"; do
  env $VAR=$v python bench.py --no-cpu-baseline --steps 3 --warmup 3 --e2e-steps 6 > gpurun_out/ab_e2e_$v.json 2> gpurun_out/ab_e2e_$v.err
  python - <<PY
import json
try:
    d=json.load(open("gpurun_out/ab_e2e_$v.json")); e=d["e2e"]
    print("$VAR=$v", "single", round(e["single_engine_ms_per_step"],3), "dual", round(e["dual_engine_ms_per_step"],3), "gather_last_chunk", round(e["stage_ms_last_chunk"]["gather"],3), "h2d", round(e["stage_ms_last_chunk"]["h2d"],2))
except Exception as ex:
    print("$v failed", ex); print(open("gpurun_out/ab_e2e_$v.err").read()[-800:])
PY
done
