#!/bin/bash
# multi-GPU bench lines: weak scaling at the driver's settings and strong scaling of fixed ensembles
N=${1:-4}
mkdir -p gpurun_out
run() { local tag=$1; shift; timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus $N "$@" > gpurun_out/r02_bench_${N}gpu_${tag}.json 2> gpurun_out/mg_${tag}.err; echo "$tag rc=$?"; tail -c 300 gpurun_out/mg_${tag}.err; }
run weak_k20 --steps 20 --warmup 5
run weak_k200 --steps 200 --warmup 5
run strong512 --steps 100 --warmup 5 --scaling strong --trajectories 512
run strong1024 --steps 100 --warmup 5 --scaling strong --trajectories 1024
python - <<PY
import json,glob
for f in sorted(glob.glob("gpurun_out/r02_bench_${N}gpu_*.json")):
    try:
        d = json.loads(open(f).read().strip().splitlines()[-1])
        print(f.split('/')[-1], d["n_gpus"], d["scaling"], d["config"]["trajectories_total"], round(d["value"]), round(d["e2e"]["value"]), d["ms_per_step"], d["allgather_ms"])
    except Exception as e:
        print(f, "unreadable", e)
PY
