#!/bin/bash
# 20 inlined job variants per sweep kernel (FCB_SWEEP_SLIM=0 build) against 5 with run-time branches (shipped), both trees
run() { echo "== $*"; env "$@" timeout 300 python tools/gpu_check.py 256 2 2>&1 | grep -E "problem setup|phase (forward|backward)|graph step|rel|FAIL|rror" ; }
run FCB_LIB=tools/bench_src/variants/lib_slim0.so
run FCB_SWEEP_NONE=1
run FCB_BALANCED=1
run FCB_LIB=tools/bench_src/variants/lib_slim0.so FCB_BALANCED=1
run FCB_BALANCED=1 FCB_LEAF=20
