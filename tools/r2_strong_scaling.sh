#!/bin/bash
# strong scaling of the fixed-width open-loop configurations (BASELINE configs[2], configs[4]) for the given rank counts
mkdir -p gpurun_out
for N in "$@"; do
  for wl in "lidcavity 1024" "pinball 512"; do
    set -- $wl; name=$1; T=$2
    out=gpurun_out/r02_strong_${name}_${N}gpu.json
    if [ "$N" = "1" ]; then
      timeout 900 python bench.py --workload $name --scaling strong --trajectories $T --steps 50 --warmup 5 > $out 2> gpurun_out/ss.err
    else
      timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29544 bench.py --gpus $N --workload $name --scaling strong --trajectories $T --steps 50 --warmup 5 > $out 2> gpurun_out/ss.err
    fi
    echo "$name N=$N rc=$?"; grep -v "OMP_NUM_THREADS\|^\*\*\*\|^$" gpurun_out/ss.err | tail -n 3
  done
done
python - <<PY
import json,glob
for f in sorted(glob.glob("gpurun_out/r02_strong_*gpu.json")):
    try:
        d = json.loads(open(f).read().strip().splitlines()[-1])
        print(f.split('/')[-1], d["n_gpus"], d["config"]["trajectories_total"], round(d["value"]), round(d["e2e"]["value"]), round(d["ms_per_step"],3), round(d["allgather_ms"],3), d["all_finite"])
    except Exception as e:
        print(f, "unreadable", e)
PY
