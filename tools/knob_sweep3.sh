#!/bin/bash
# re-sweep of the element-patch size and the ring geometry of the sweeps on the depth-bounded tree (cylinder, 256 trajectories)
run() { echo "== $*"; env "$@" timeout 300 python tools/gpu_check.py 256 2 2>&1 | grep -E "phase (forward|backward|element)|graph step|FAIL|rror" ; }
run FCB_NOP=1
run FCB_PATCH_CELLS=20
run FCB_PATCH_CELLS=28
run FCB_PATCH_CELLS=32
run FCB_SWEEP_SLOTS=48,36,24,12
run FCB_SWEEP_SLOTS=48,36,12,24
run FCB_SWEEP_KSLOTS=12
run FCB_SWEEP_KSLOTS=36
run FCB_SWEEP_WANT=1.5
run FCB_SWEEP_WANT=3.0
run FCB_SWEEP_MAXWARPS=8
run FCB_SWEEP_MAXWARPS=2
