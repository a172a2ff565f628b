#!/bin/bash
# One N-GPU lease, trimmed for the budget: the box's host<->device ceiling with all N ranks copying at once, then the default
# N-rank bench line (weak scaling, e2e with its own pcie_ceiling, and the attached fixed-workload strong_scaling figure).
# Usage (under gpurun --gpus N): bash tools/gpu_multi_n.sh <tag> <N>
TAG=${1:-multi}; N=${2:-8}; OUT=gpurun_out/$TAG; mkdir -p $OUT
nvidia-smi topo -m > $OUT/topo.txt 2>&1; nproc >> $OUT/topo.txt; free -g >> $OUT/topo.txt
TR() { n=$1; shift; python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port $((29500 + n)) "$@"; }
TR $N bench.py --stage pcie --mb 512 --reps 20 2>$OUT/pcie_$N.err | tail -1 | tee -a $OUT/pcie.jsonl
TR $N bench.py --gpus $N --steps 10 --warmup 3 --no-cpu-baseline 2>$OUT/weak_$N.err | tail -1 | tee -a $OUT/weak.jsonl | python -c "
import json,sys
d=json.loads(sys.stdin.read()); e=d['e2e']; s=d.get('strong_scaling') or {}
print('weak N', d['n_gpus'], 'value', d['value'], 'ms', d['ms_per_step'], '| e2e', e['value'], 'ms', e['ms_per_step'], e['coords_mode'], 'engines', e['engines_per_gpu'], 'single', e['single_engine_ms_per_step'], 'dual', e['dual_engine_ms_per_step'])
print('ceiling', e.get('pcie_ceiling'), 'achieved/rank', e.get('h2d_gbs_per_rank_achieved'), 'frac', e.get('h2d_fraction_of_ceiling'))
print('strong', {k: s.get(k) for k in ('n_gpus','ms_per_step','value','identical_to_single_gpu','launches_per_step')})"
tail -3 $OUT/weak_$N.err
