#!/bin/bash
mkdir -p gpurun_out
run() { echo "== $*"; env "$@" python tools/gpu_check.py 256 20 2>&1 | grep -E "clusters:|worst rel|final field|phase (forward|backward)|graph step|rror" | sort -u ; }
run FCB_VERBOSE=1 FCB_CLUSTER_ROWS=512 FCB_CLUSTER_HEIGHT=3
run FCB_VERBOSE=1 FCB_CLUSTER_ROWS=512 FCB_CLUSTER_HEIGHT=6
run FCB_VERBOSE=1 FCB_CLUSTER_ROWS=400 FCB_CLUSTER_HEIGHT=6
FCB_CLUSTER_ROWS=512 FCB_CLUSTER_HEIGHT=6 timeout 300 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -k regex:"k_cluster_sweep|k_front_sweep" -c 60 --csv --log-file gpurun_out/cl_launches.csv python tools/gpu_check.py 256 2 > gpurun_out/ncu_cl.log 2>&1
python - <<'PY'
import csv
lines=open('gpurun_out/cl_launches.csv').read().splitlines()
st=[i for i,l in enumerate(lines) if l.startswith('"ID"')][0]
ks={}
for r in csv.DictReader(lines[st:]):
    d=ks.setdefault(r['ID'],{'name':r['Kernel Name'][:40],'grid':r.get('Grid Size')})
    d[r['Metric Name']]=float(r['Metric Value'].replace(',',''))
for j in list(ks)[-22:]:
    d=ks[j]; print(d['name'].ljust(40), d['grid'], '%.1f us'%(d['gpu__time_duration.sum']/1e3), '%.1f MB'%((d.get('dram__bytes_read.sum',0)+d.get('dram__bytes_write.sum',0))/1e6))
PY
