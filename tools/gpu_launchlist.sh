#!/bin/bash
# per-kernel durations (ncu, no replay metrics) of the 64-frame batch: gpu_launchlist.sh <tag> [kernel regex]
T=$1; K=${2:-.}
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"$K" --csv --log-file gpurun_out/${T}_ll.csv python tools/prof_kernels.py 64 ${UPR:-1} > /dev/null 2>&1
python - <<PY
import csv, collections
rows = list(csv.reader(open('gpurun_out/${T}_ll.csv')))
hi = next(i for i,r in enumerate(rows) if 'Kernel Name' in r)
hdr = rows[hi]; kn = hdr.index('Kernel Name'); mv = hdr.index('Metric Value')
agg = collections.OrderedDict()
for r in rows[hi+1:]:
    if len(r) > mv: agg.setdefault(r[kn].split('(')[0], []).append(float(r[mv]))
for k, v in agg.items(): print(f"{k:50s} n={len(v)} min {min(v)/1e3:9.2f} us  = {min(v)/64e3:6.2f} us/frame")
PY
