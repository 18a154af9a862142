#!/bin/bash
# one GPU-box visit: parity tests, smoke, bench, ncu launch list.  Outputs -> gpurun_out/
set -u
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/smi.txt 2>&1
python -c "import __graft_entry__ as g; g.build()" > gpurun_out/build.log 2>&1
timeout 900 python -m pytest tests -m gpu -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
tail -40 gpurun_out/pytest_gpu.log
timeout 300 python __graft_entry__.py smoke > gpurun_out/smoke.log 2>&1; echo "smoke rc=$?" >> gpurun_out/smoke.log
tail -3 gpurun_out/smoke.log
timeout 600 python bench.py --steps 30 --warmup 5 > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "bench rc=$?"
tail -c 3500 gpurun_out/bench.json; tail -5 gpurun_out/bench.err
timeout 600 python bench.py --steps 30 --warmup 5 --no-sort --no-cpu > gpurun_out/bench_nosort.json 2>> gpurun_out/bench.err; echo "bench nosort rc=$?"
tail -c 3500 gpurun_out/bench_nosort.json
if [ "${NCU:-1}" = "1" ]; then
  timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches.csv \
    python bench.py --steps 3 --warmup 3 --no-cpu > gpurun_out/bench_ncu.log 2>&1
  echo "ncu rc=$?"
fi
