#!/bin/bash
# Round-2 first GPU pass: whole GPU suite, bench line, sweep timeline.
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q --deselect tests/test_gpu_configs.py::test_config2_pinball_re100_rotation_512 > gpurun_out/r02_gpu_tests_a.log 2>&1
echo "pytest rc=$?"; tail -n 5 gpurun_out/r02_gpu_tests_a.log
timeout 400 python bench.py > gpurun_out/r02_bench_a.json 2> gpurun_out/bench_a.err; echo "bench rc=$?"; tail -c 600 gpurun_out/bench_a.err
FCB_SWEEP_DEBUG=gpurun_out/sweep_dbg.bin timeout 300 python tools/gpu_check.py 256 4 > gpurun_out/gpu_check_a.log 2>&1
python tools/sweep_timeline.py gpurun_out/sweep_dbg.bin > gpurun_out/r02_sweep_timeline_a.txt 2>&1
cat gpurun_out/r02_sweep_timeline_a.txt
