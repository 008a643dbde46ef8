#!/bin/bash
# in-place leaves (FCB_LEAF_INPLACE=1: no y rows and no solver-order x rows for the leaves) against the plain plan
run() { local B=$1; shift; echo "== B=$B $*"; env "$@" timeout 300 python tools/gpu_check.py $B 2 2>&1 | grep -E "worst rel|phase (forward|backward)|graph step|FAIL|rror" ; }
run 256 FCB_LEAF_INPLACE=0
run 256 FCB_LEAF_INPLACE=1
run 256 FCB_LEAF_INPLACE=0
run 256 FCB_LEAF_INPLACE=1
run 72 FCB_LEAF_INPLACE=1
