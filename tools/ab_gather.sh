# A/B of the N-GPU bench variants (scratch tool): bash tools/ab_gather.sh [N]
N=${1:-2}
run() { tag=$1; shift; env "$@" python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 10 --warmup 3 --e2e-steps 2 $EXTRA > gpurun_out/ab_$tag.json 2> gpurun_out/ab_$tag.err; python - <<PY
import json
try:
    d=json.loads(open("gpurun_out/ab_$tag.json").read().strip().splitlines()[-1])
    print("$tag", round(d["ms_per_step"],4), {k:round(v,3) for k,v in d["roofline"]["stage_ms_per_step"].items()}, "e2e", round(d["e2e"]["ms_per_step"],2), d.get("per_rank"))
except Exception as e:
    print("$tag failed", e); print(open("gpurun_out/ab_$tag.err").read()[-1500:])
PY
}
EXTRA="--gather final" run final A=1

