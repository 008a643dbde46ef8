#!/bin/bash
# Round measurements on one B200 (run under gpurun): GPU suite, bench lines (no profiler), then ONE ncu pass (launch list with
# per-launch time and DRAM bytes) of the same bench command.  Results land in gpurun_out/; copy what should be judged to profiles/.
R=${1:-r02}
mkdir -p gpurun_out
timeout 1700 python -m pytest tests -m gpu -x -q > gpurun_out/${R}_gpu_tests_final.log 2>&1; echo "pytest rc=$?"; tail -n 4 gpurun_out/${R}_gpu_tests_final.log
timeout 400 python bench.py > gpurun_out/${R}_bench_final.json 2> gpurun_out/bench.err || exit 1
timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/${R}_bench_k20.json 2>> gpurun_out/bench.err
timeout 300 python bench.py --time-scheme cn --no-cpu-baseline > gpurun_out/${R}_bench_cn.json 2>> gpurun_out/bench.err
timeout 300 python bench.py --impl reference --steps 20 --warmup 5 > gpurun_out/${R}_bench_reference.json 2>> gpurun_out/bench.err
timeout 400 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -c 400 --csv \
    --log-file gpurun_out/${R}_launches_bench.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/ncu.log 2>&1
python tools/kernel_summary.py gpurun_out/${R}_launches_bench.csv gpurun_out/${R}_kernel_summary.json
python - <<PY
import json
for f in ("${R}_bench_final", "${R}_bench_k20", "${R}_bench_cn", "${R}_bench_reference"):
    try:
        d = json.loads(open(f"gpurun_out/{f}.json").read().strip().splitlines()[-1])
        print(f, d.get("value"), (d.get("e2e") or {}).get("value"), (d.get("roofline") or {}).get("frac"), d.get("allgather_ms"), (d.get("clocks") or {}), (d.get("cpu_baseline") or {}).get("value"))
    except Exception as e:
        print(f, "unreadable", e)
PY
