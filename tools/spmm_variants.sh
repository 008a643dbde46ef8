#!/bin/bash
# k_spmm_mma tuning builds (staged columns per block / CTAs per SM): Crank-Nicolson cylinder, B = 256
run() { echo "== $*"; env "$@" FCB_SCHEME=cn python tools/gpu_check.py 256 4 2>&1 | grep -E "k_spmm|phase spmm|worst rel|rror" ; }
run FCB_NOP=1
for f in tools/bench_src/variants/libfcb200_spmm_*.so; do run FCB_LIB=$f; done
