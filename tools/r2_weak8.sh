#!/bin/bash
mkdir -p gpurun_out
N=${1:-8}
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29555 bench.py --gpus $N --steps 20 --warmup 5 > gpurun_out/r02_bench_${N}gpu_weak_k20.json 2> gpurun_out/w8.err; echo "rc=$?"
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29556 bench.py --impl reference --gpus $N --steps 20 --warmup 5 > gpurun_out/r02_bench_${N}gpu_reference.json 2>> gpurun_out/w8.err; echo "rc=$?"
python - <<PY
import json
for f in ("r02_bench_${N}gpu_weak_k20", "r02_bench_${N}gpu_reference"):
    d = json.loads(open(f"gpurun_out/{f}.json").read().strip().splitlines()[-1])
    print(f, d.get("n_gpus"), round(d["value"]), round(d["e2e"]["value"]), d.get("ms_per_step"), d.get("allgather_ms"), (d.get("cpu_baseline") or {}).get("cores"))
PY
