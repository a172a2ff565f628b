#!/bin/bash
# One multi-GPU lease: host->device ceiling at N = 1, 2, 4, 8 ranks; the default bench at N; the fixed-workload (strong) mode at N = 1..8
# Usage (under gpurun --gpus 8): bash tools/gpu_multi.sh <tag> <maxN>
TAG=${1:-multi}; MAXN=${2:-8}; OUT=gpurun_out/$TAG; mkdir -p $OUT
nvidia-smi topo -m > $OUT/topo.txt 2>&1; nproc >> $OUT/topo.txt; free -g >> $OUT/topo.txt
TR() { n=$1; shift; if [ $n -eq 1 ]; then python "$@"; else python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port $((29500 + n)) "$@"; fi; }
for n in 1 2 4 8; do [ $n -le $MAXN ] && TR $n bench.py --stage pcie --mb 512 --reps 20 2>$OUT/pcie_$n.err | tail -1 | tee -a $OUT/pcie.jsonl; done
for n in 1 2 4 8; do [ $n -le $MAXN ] && TR $n bench.py --gpus $n --scaling strong --steps 20 --warmup 3 2>$OUT/strong_$n.err | tail -1 | tee -a $OUT/strong.jsonl | head -c 700; echo; done
WEAK="$MAXN"; [ $MAXN -gt 4 ] && WEAK="$MAXN 4"
for n in $WEAK; do [ $n -le $MAXN ] && [ $n -gt 1 ] && TR $n bench.py --gpus $n --steps 10 --warmup 3 --no-strong --no-cpu-baseline 2>$OUT/weak_$n.err | tail -1 | tee -a $OUT/weak.jsonl | python -c "
import json,sys
d=json.loads(sys.stdin.read()); e=d['e2e']; print('weak N', d['n_gpus'], 'ms', d['ms_per_step'], 'e2e ms', e['ms_per_step'], e['coords_mode'], e['coords_mode_ms_single_engine'], 'single', e['single_engine_ms_per_step'], 'dual', e['dual_engine_ms_per_step'], 'ceiling', e.get('pcie_ceiling'), 'frac', e.get('h2d_fraction_of_ceiling'))"; done
