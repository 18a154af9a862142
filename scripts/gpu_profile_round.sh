#!/bin/bash
# gpurun driver: GPU tests, ncu --set full captures of the step kernels (one launch each, inside scripts/step_ab.py),
# the ncu launch list of a short bench run, and the bench line itself
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -x > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log; tail -3 gpurun_out/pytest_gpu.log
for k in ${KERNELS:-k_step_a k_step_bw k_step_nnq}; do
  KERNEL=$k SKIP=30 TAG=$k bash scripts/gpu_prof.sh
done
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches.csv python bench.py --steps 2 --warmup 3 --no-cpu --no-converged > gpurun_out/bench_ncu.log 2>&1
echo "launch list rc=$?"
