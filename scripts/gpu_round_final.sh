#!/bin/bash
# last visit of the round: smoke, GPU suite, bench lines (ours + reference arm), ncu --set full of the step kernels and of the
# TCN pair GEMM, ncu launch lists (bench, TCN), cotter-pin stand-in, TCN timing probe
mkdir -p gpurun_out
timeout 300 python __graft_entry__.py smoke > gpurun_out/smoke.log 2>&1; echo "smoke rc=$?" >> gpurun_out/smoke.log; tail -2 gpurun_out/smoke.log
timeout 900 python -m pytest tests -m gpu -q -x > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log; tail -3 gpurun_out/pytest_gpu.log
timeout 900 python bench.py > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "bench rc=$?"; tail -3 gpurun_out/bench.err
timeout 900 python bench.py --impl reference > gpurun_out/bench_reference.json 2>> gpurun_out/bench.err; echo "reference rc=$?"
for k in k_step_a k_step_bw k_step_nnq k_step_meshq k_step_meshq2; do
  KERNEL="^$k\$" SKIP=30 TAG=$k bash scripts/gpu_prof.sh
done
TCN_NO_KINETO=1 timeout 300 ncu --set full --clock-control none --import-source on -k regex:k_tcn_pair_mma -s 49 -c 1 -f \
  -o gpurun_out/prof_k_tcn_pair_mma python scripts/tcn_prof.py > gpurun_out/ncu_k_tcn_pair_mma.log 2>&1; echo "ncu tcn rc=$?"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/launches.csv \
    python bench.py --steps 30 --warmup 5 --no-cpu --no-converged > gpurun_out/bench_ncu.log 2>&1; echo "ncu launches rc=$?"
TCN_NO_KINETO=1 MIDAS_B200_TCN_NO_PDL=1 timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/tcn_launches.csv \
  python scripts/tcn_prof.py > gpurun_out/tcn_ncu.log 2>&1; echo "tcn launch list rc=$?"
timeout 200 python scripts/tcn_prof.py > gpurun_out/tcn_prof.json 2> gpurun_out/tcn_prof.err
: > gpurun_out/pin.json
timeout 300 python scripts/pin_ab.py 2> gpurun_out/pin.err | tee -a gpurun_out/pin.json
AB_GRAPH=1 timeout 300 python scripts/pin_ab.py 2>> gpurun_out/pin.err | tee -a gpurun_out/pin.json
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,clocks_throttle_reasons.active --format=csv > gpurun_out/smi.txt
