#!/bin/bash
# A/B of prebuilt library variants (variants/*.so, built here with different -D flags): bench.py kernel times
mkdir -p gpurun_out
if [ "${TESTS:-0}" = "1" ]; then
timeout 900 python -m pytest tests -m gpu -q -x > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
tail -5 gpurun_out/pytest_gpu.log
fi
for lib in variants/*.so; do
  for rep in 1 2; do
  MIDAS_B200_LIB=$PWD/$lib timeout 300 python bench.py --steps ${STEPS:-100} --warmup 5 --no-cpu ${BENCH_ARGS:-} > gpurun_out/v.json 2> gpurun_out/v.err
  python - "$lib" <<'PY'
import json,sys
try:
    d=json.load(open('gpurun_out/v.json'))
    k=d['roofline']['sweep']['kernels_ms']
    print(sys.argv[1], 'ms/step %.4f'%d['ms_per_step'], ' '.join('%.1f'%(1e3*v) for v in k.values()), 'e2e %.3g'%d['e2e']['value'])
except Exception as e:
    print(sys.argv[1], 'FAILED', e); print(open('gpurun_out/v.err').read()[-800:])
PY
  done
done
