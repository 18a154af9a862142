#!/bin/bash
# per-launch device times (ncu, serialised) of the step kernels inside scripts/step_ab.py
mkdir -p gpurun_out
AB_STEPS=${AB_STEPS:-30} timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"k_step|k_codebook_query" -s 100 -c 200 --csv --log-file gpurun_out/launches_ab.csv python scripts/step_ab.py > gpurun_out/launches_ab.log 2>&1
python - <<'PY'
import csv,collections
rows=[r for r in csv.reader(open('gpurun_out/launches_ab.csv')) if len(r)>10]
h=rows[0]; ki=h.index('Kernel Name'); vi=h.index('Metric Value')
agg=collections.defaultdict(list)
for r in rows[1:]:
    try: agg[r[ki].split('(')[0]].append(float(r[vi].replace(',','')))
    except: pass
for k,v in agg.items(): print(k, len(v), 'avg us %.1f'%(sum(v)/len(v)/1e3 if max(v)>1e4 else sum(v)/len(v)), 'max %.1f'%(max(v)/1e3 if max(v)>1e4 else max(v)))
PY
