#!/bin/bash
# full GPU visit: tests, smoke, bench (+no-sort, +reference arm), launch list, ncu --set full of the step kernels
set -u
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.build()" > gpurun_out/build.log 2>&1
timeout 900 python -m pytest tests -m gpu -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
tail -5 gpurun_out/pytest_gpu.log
timeout 300 python __graft_entry__.py smoke > gpurun_out/smoke.log 2>&1; echo "smoke rc=$?" >> gpurun_out/smoke.log; tail -2 gpurun_out/smoke.log
nvidia-smi --query-gpu=index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap --format=csv -lms 200 > gpurun_out/clocks.csv &
SMI=$!
timeout 900 python bench.py > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "bench rc=$?"; tail -c 3800 gpurun_out/bench.json; tail -3 gpurun_out/bench.err
kill $SMI
timeout 600 python bench.py --no-sort --no-cpu > gpurun_out/bench_nosort.json 2>> gpurun_out/bench.err; echo "bench nosort rc=$?"
timeout 600 python bench.py --impl reference --steps 20 --warmup 2 > gpurun_out/bench_reference.json 2>> gpurun_out/bench.err; echo "reference rc=$?"; tail -c 1500 gpurun_out/bench_reference.json
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/launches.csv \
    python bench.py --steps 30 --warmup 5 --no-cpu > gpurun_out/bench_ncu.log 2>&1; echo "ncu launches rc=$?"
for k in k_step_a k_step_nnq k_step_bw; do
  timeout 900 ncu --set full --clock-control none --import-source on -k regex:$k -s 30 -c 1 -f -o gpurun_out/prof_$k \
     python bench.py --steps 4 --warmup 30 --no-cpu > gpurun_out/ncu_$k.log 2>&1
  echo "$k rc=$?"
done
