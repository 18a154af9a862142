#!/bin/bash
# gpurun driver: [TESTS=1: the whole GPU suite], TCN timing probe, ncu launch list of TCN forwards, ncu --set full of the
# transposed convolution's pair GEMM (the 10th k_tcn_pair_mma launch of a forward)
mkdir -p gpurun_out
if [ "${TESTS:-1}" = "1" ]; then
timeout 900 python -m pytest tests -m gpu -q -x > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log; tail -3 gpurun_out/pytest_gpu.log
fi
timeout 200 python scripts/tcn_prof.py > gpurun_out/tcn_prof.json 2> gpurun_out/tcn_prof.err; tail -2 gpurun_out/tcn_prof.err
TCN_NO_KINETO=1 timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/tcn_launches.csv \
  python scripts/tcn_prof.py > gpurun_out/tcn_ncu.log 2>&1; echo "launch list rc=$?"
TCN_NO_KINETO=1 timeout 300 ncu --set full --clock-control none --import-source on -k regex:k_tcn_pair_mma -s ${SKIP:-49} -c 1 -f \
  -o gpurun_out/prof_k_tcn_pair_mma python scripts/tcn_prof.py > gpurun_out/ncu_k_tcn_pair_mma.log 2>&1; echo "ncu full rc=$?"
