#!/bin/bash
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.build()" > gpurun_out/build.log 2>&1
if [ "${TESTS:-1}" = "1" ]; then
timeout 900 python -m pytest tests -m gpu -q -x > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
tail -15 gpurun_out/pytest_gpu.log
fi
show() { python - "$1" <<'PY'
import json,sys
d=json.load(open(sys.argv[1]))
for k,v in d.items(): print(k, v['us_all'], v['fallbacks_all'], v['on_surface'])
PY
}
python scripts/variants.py > gpurun_out/variants.json 2> gpurun_out/variants.err; show gpurun_out/variants.json
tail -3 gpurun_out/variants.err
for v in ${VARIANTS:-}; do
  nvcc -gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 -shared -Xcompiler -fPIC $v -o /tmp/libvar.so midastouch_b200/csrc/midas_b200.cu
  echo "variant $v"; MIDAS_B200_LIB=/tmp/libvar.so python scripts/variants.py > gpurun_out/variants_v.json 2>> gpurun_out/variants.err; show gpurun_out/variants_v.json
done
for k in ${KERNELS:-}; do
  timeout 900 ncu --set full --clock-control none --import-source on -k regex:$k -s ${SKIP:-8} -c 1 -f -o gpurun_out/prof_$k \
     python bench.py --steps 4 --warmup ${WARM:-8} --no-cpu > gpurun_out/ncu_$k.log 2>&1
  echo "$k rc=$?"
done
