#!/bin/bash
run() { echo "== B=$1 $2"; env $2 python tools/gpu_check.py $1 4 2>&1 | grep -E "phase (forward|backward|element)|graph step|rror" | sort -u ; }
for B in 1 64 128; do
  run $B FCB_CLUSTER_ROWS=0
  run $B "FCB_CLUSTER_ROWS=512 FCB_CLUSTER_HEIGHT=3"
  run $B "FCB_CLUSTER_ROWS=512 FCB_CLUSTER_HEIGHT=6"
done
