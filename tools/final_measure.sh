#!/bin/bash
# Round-end measurements on one B200 (run under gpurun): bench lines first (no profiler), then ONE ncu pass (launch list with
# per-launch time and DRAM bytes) of the same bench command.  Results land in gpurun_out/; copy what should be judged to profiles/.
mkdir -p gpurun_out
timeout 400 python bench.py > gpurun_out/r01_bench_final.json 2> gpurun_out/bench.err || exit 1
timeout 300 python bench.py --time-scheme cn --no-cpu-baseline > gpurun_out/r01_bench_cn.json 2>> gpurun_out/bench.err || exit 1
timeout 300 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/r01_bench_reference.json 2>> gpurun_out/bench.err
timeout 400 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -c 400 --csv \
    --log-file gpurun_out/launches.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/ncu.log 2>&1
python - <<'PY'
import json
for f in ("r01_bench_final", "r01_bench_cn", "r01_bench_reference"):
    try:
        d = json.loads(open(f"gpurun_out/{f}.json").read().strip().splitlines()[-1])
        print(f, d.get("value"), (d.get("e2e") or {}).get("value"), (d.get("roofline") or {}).get("frac"), (d.get("clocks") or {}))
    except Exception as e:
        print(f, "unreadable", e)
PY
