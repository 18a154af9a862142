#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -x > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
tail -15 gpurun_out/pytest_gpu.log
timeout 300 python bench.py --steps 100 --warmup 5 --no-cpu > gpurun_out/v.json 2> gpurun_out/v.err
python - <<'PY'
import json
d=json.load(open('gpurun_out/v.json'))
k=d['roofline']['sweep']['kernels_ms']
print('ms/step %.4f'%d['ms_per_step'], ' '.join('%.1f'%(1e3*v) for v in k.values()), 'e2e %.3g'%d['e2e']['value'], d['engine_stats'])
PY
timeout 600 python scripts/sweeps.py 2>&1 | tail -8
