#!/bin/bash
# k_element_patch tuning builds: register prefetch on/off, CTAs per SM, cells per patch (cylinder, B = 256)
run() { echo "== $*"; env "$@" python tools/gpu_check.py 256 4 2>&1 | grep -E "phase element|graph step|rror" ; }
run FCB_NOP=1
run FCB_LIB=tools/bench_src/variants/libfcb200_ep_2_0.so
run FCB_LIB=tools/bench_src/variants/libfcb200_ep_3_0.so FCB_PATCH_CELLS=16
run FCB_LIB=tools/bench_src/variants/libfcb200_ep_3_0.so FCB_PATCH_CELLS=14
run FCB_LIB=tools/bench_src/variants/libfcb200_ep_3_0.so FCB_PATCH_CELLS=24
