#!/bin/bash
run() { echo "== $*"; env "$@" python tools/gpu_check.py 256 20 2>&1 | grep -E "problem setup|worst rel|final field|phase (forward|backward)|graph step|rror" | sort -u ; }
run FCB_AMALGAMATE=0
run FCB_AMALGAMATE=4
run FCB_AMALGAMATE=3
run FCB_AMALGAMATE=5
run FCB_AMALGAMATE=2
