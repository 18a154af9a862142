#!/bin/bash
mkdir -p gpurun_out
for rep in 1 2; do
timeout 300 python bench.py --steps 100 --warmup 5 --no-cpu > gpurun_out/v.json 2> gpurun_out/v.err
python - <<'PY'
import json
d=json.load(open('gpurun_out/v.json'))
k=d['roofline']['sweep']['kernels_ms']
print('ms/step %.4f'%d['ms_per_step'], ' '.join('%.1f'%(1e3*v) for v in k.values()), 'e2e %.3g'%d['e2e']['value'], d['engine_stats'])
PY
done
