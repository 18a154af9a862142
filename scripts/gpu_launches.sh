#!/bin/bash
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.build()" > gpurun_out/build.log 2>&1
if [ "${TESTS:-1}" = "1" ]; then
timeout 900 python -m pytest tests -m gpu -q -x > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
tail -5 gpurun_out/pytest_gpu.log
fi
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/launches.csv \
    python bench.py --steps 30 --warmup 5 --no-cpu > gpurun_out/bench_ncu.log 2>&1
echo "ncu rc=$?"
timeout 600 python bench.py --steps 30 --warmup 5 --no-cpu > gpurun_out/bench.json 2> gpurun_out/bench.err; tail -c 1200 gpurun_out/bench.json
for k in ${KERNELS:-}; do
  timeout 900 ncu --set full --clock-control none --import-source on -k regex:$k -s ${SKIP:-24} -c 1 -f -o gpurun_out/prof_$k \
     python bench.py --steps 4 --warmup ${WARM:-25} --no-cpu > gpurun_out/ncu_$k.log 2>&1
  echo "$k rc=$?"
done
