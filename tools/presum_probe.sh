#!/bin/bash
# pre-summed forward gathers (FCB_PRESUM=h: fronts of height >= h read one plane written by a gather-sum) against three-plane gathers
run() { echo "== $*"; env "$@" timeout 300 python tools/gpu_check.py 256 2 2>&1 | grep -E "problem setup|phase (forward|backward)|graph step|worst rel|FAIL|rror" ; }
run FCB_PRESUM=0
run FCB_PRESUM=1
run FCB_PRESUM=2
run FCB_PRESUM=3
run FCB_PRESUM=4
run FCB_PRESUM=6
run FCB_PRESUM=2 FCB_BALANCED=0
