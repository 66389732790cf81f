timeout 100 python tools/prof_kernels.py 1 1 && timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none --cache-control none --csv --log-file gpurun_out/r2_ll_b1.csv python tools/prof_kernels.py 1 1 > /dev/null 2>&1
python - <<PY
import csv, collections
rows = list(csv.reader(open('gpurun_out/r2_ll_b1.csv')))
hi = next(i for i,r in enumerate(rows) if 'Kernel Name' in r)
hdr = rows[hi]; kn = hdr.index('Kernel Name'); mv = hdr.index('Metric Value')
agg = collections.OrderedDict()
for r in rows[hi+1:]:
    if len(r) > mv: agg.setdefault(r[kn].split('(')[0], []).append(float(r[mv]))
for k, v in agg.items(): print(f"{k:50s} n={len(v)} min {min(v)/1e3:9.2f} us")
PY
