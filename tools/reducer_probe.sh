#!/bin/bash
# which consumer warp of a k-split CTA waits for the seed rows and reduces: warp 0 (-DFCB_KS_REDUCER=0 build) or warp 3 (shipped)
run() { local B=$1; shift; echo "== B=$B $*"; env "$@" timeout 300 python tools/gpu_check.py $B 2 2>&1 | grep -E "worst rel|phase (forward|backward)|graph step|FAIL|rror" ; }
for B in 32 256; do
run $B FCB_LIB=tools/bench_src/variants/lib_rw0.so FCB_CLUSTER_ROWS=0
run $B FCB_CLUSTER_ROWS=0
run $B FCB_LIB=tools/bench_src/variants/lib_rw0.so FCB_CLUSTER_ROWS=0
run $B FCB_CLUSTER_ROWS=0
done
