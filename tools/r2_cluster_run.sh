#!/bin/bash
mkdir -p gpurun_out
FCB_VERBOSE=1 timeout 300 python tools/gpu_check.py 256 20 > gpurun_out/gpu_check_b.log 2>&1; echo "gpu_check rc=$?"; grep -E "clusters|worst|final field|spread|phase|total|graph step" gpurun_out/gpu_check_b.log
timeout 900 python -m pytest tests/test_gpu_parity.py -x -q -k "widths or golden or closed_loop or restart" > gpurun_out/r02_gpu_tests_b.log 2>&1; echo "pytest rc=$?"; tail -n 15 gpurun_out/r02_gpu_tests_b.log
