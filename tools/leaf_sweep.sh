#!/bin/bash
run() { echo "== $*"; env "$@" python tools/gpu_check.py 256 2 2>&1 | grep -E "problem setup|phase (forward|backward)|graph step" ; }
run FCB_LEAF=12
run FCB_LEAF=24
run FCB_LEAF=32
run FCB_LEAF=16 FCB_TOP=3
