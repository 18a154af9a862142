#!/bin/bash
# round-2 kernels under compute-sanitizer: the fused step (flat k_step_bw, team search in k_step_nnq, search records),
# the k-nearest query and the long neighbour lists
mkdir -p gpurun_out
SEL='fused_step_vs_oracle or with_prune or all_drifted or philox_teacher or k_neighbours or r3_se3_other or low_var_sizes or heavy_fallback'
timeout 2400 compute-sanitizer --tool memcheck --error-exitcode 7 --print-limit 20 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "$SEL" > gpurun_out/memcheck3.log 2>&1; echo "memcheck3 rc=$?"
grep -a "ERROR SUMMARY\|passed\|failed\|Invalid\|out of bounds" gpurun_out/memcheck3.log | tail -8
timeout 1800 compute-sanitizer --tool racecheck --error-exitcode 7 --print-limit 20 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "fused_step_vs_oracle or with_prune or low_var_heavy" > gpurun_out/racecheck3.log 2>&1; echo "racecheck3 rc=$?"
grep -a "RACECHECK SUMMARY\|passed\|failed\|hazard" gpurun_out/racecheck3.log | tail -8
