#!/bin/bash
# sweep tuning knobs of k_front_sweep: prints forward/backward phase ms per setting
mkdir -p gpurun_out
run() { echo "== $*"; env "$@" python tools/gpu_check.py 256 2 2>&1 | grep -E "phase (forward|backward)|graph step" ; }
run FCB_NOP=1
run FCB_SWEEP_SLOTS=48,36,24,12
run FCB_SWEEP_SLOTS=48,36,36,12
run FCB_SWEEP_SLOTS=96,72,36,12
run FCB_SWEEP_WANT=1.0
run FCB_SWEEP_WANT=3.0
run FCB_SWEEP_WANT=1.0 FCB_SWEEP_SLOTS=48,36,36,12
run FCB_SWEEP_KSLOTS=48
