#!/bin/bash
# depth-bounded dissection (FCB_BALANCED=1) against the free dissection: setup line, oracle check, phase times, step time
run() { echo "== $*"; env "$@" timeout 300 python tools/gpu_check.py 256 2 2>&1 | grep -E "problem setup|max rel|phase (forward|backward|element)|graph step|FAIL|Error|error" ; }
run FCB_BALANCED=0
run FCB_BALANCED=1
run FCB_BALANCED=1 FCB_LEAF=20
run FCB_BALANCED=1 FCB_TOP=1
run FCB_BALANCED=1 FCB_TOP=3
run FCB_BALANCED=1 FCB_LEAF=32
