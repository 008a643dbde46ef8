#!/bin/bash
# narrow ensembles (the per-GPU share of a strong-scaling run): tiling targets and the subtree-cluster plan at 64 / 128 trajectories
run() { local B=$1; shift; echo "== B=$B $*"; env "$@" timeout 300 python tools/gpu_check.py $B 2 2>&1 | grep -E "phase (forward|backward)|graph step|FAIL|rror" ; }
for B in 64 128; do
run $B FCB_NOP=1
run $B FCB_CLUSTER_ROWS=512 FCB_CLUSTER_HEIGHT=3
run $B FCB_SWEEP_WANT=1.0
run $B FCB_SWEEP_WANT=3.0
run $B FCB_SWEEP_WANT_FWD=2.0
done
