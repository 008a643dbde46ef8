#!/bin/bash
run() { echo "== $*"; env "$@" python tools/gpu_check.py 256 2 2>&1 | grep -E "phase (forward|backward)|graph step" ; }
run FCB_SWEEP_KSLOTS=12
run FCB_SWEEP_KSLOTS=36
run FCB_SWEEP_KSPLIT=0
run FCB_SWEEP_WANT=1.5
