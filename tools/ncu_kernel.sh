#!/bin/bash
# One `ncu --set full` capture of a single kernel launch inside tools/gpu_check.py (run under gpurun, AFTER the same
# command has exited 0 without ncu):   bash tools/ncu_kernel.sh <kernel regex> <launches to skip> <report name> [VAR=value ...]
#   bash tools/ncu_kernel.sh k_element_patch 6 r01_element
#   bash tools/ncu_kernel.sh k_spmm_mma 3 r01_spmm_mma FCB_SCHEME=cn
# Read it back with tools/ncu_hot.py gpurun_out/<name>.ncu-rep 0  and  ncu -i ... --page raw --csv.
set -u
regex=$1; skip=$2; name=$3; shift 3
mkdir -p gpurun_out
env "$@" timeout 280 ncu --set full --import-source on --clock-control none -k "regex:$regex" -s "$skip" -c 1 \
    -o "gpurun_out/$name" python tools/gpu_check.py 256 2 > "gpurun_out/ncu_$name.log" 2>&1
tail -2 "gpurun_out/ncu_$name.log"
ls -la gpurun_out/"$name".ncu-rep
