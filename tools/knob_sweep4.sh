#!/bin/bash
# CTAs-per-SM target of the forward launches alone (the backward launches keep 2.0)
run() { echo "== $*"; env "$@" timeout 300 python tools/gpu_check.py 256 2 2>&1 | grep -E "phase (forward|backward)|graph step|FAIL|rror" ; }
run FCB_SWEEP_WANT_FWD=0.5
run FCB_SWEEP_WANT_FWD=0.75
run FCB_SWEEP_WANT_FWD=1.0
run FCB_SWEEP_WANT_FWD=1.25
run FCB_SWEEP_WANT_FWD=1.25 FCB_SWEEP_WANT=1.75
run FCB_SWEEP_WANT_FWD=1.25 FCB_SWEEP_WANT=2.25
