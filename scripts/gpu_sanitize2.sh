#!/bin/bash
mkdir -p gpurun_out
SEL='fused_step_vs_oracle or philox_teacher or spatial_sort or select_k or stale_hints or se3_nn_dropin'
timeout 2400 compute-sanitizer --tool memcheck --error-exitcode 7 --print-limit 20 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "$SEL" > gpurun_out/memcheck2.log 2>&1; echo "memcheck2 rc=$?"
grep -a "ERROR SUMMARY\|passed\|failed\|Invalid\|out of bounds" gpurun_out/memcheck2.log | tail -8
timeout 1200 compute-sanitizer --tool racecheck --error-exitcode 7 --print-limit 20 python -m pytest tests/test_gpu_parity.py tests/test_gpu_tcn.py -m gpu -q -x -k "fused_step_vs_oracle or level_counts or query_batched" > gpurun_out/racecheck2.log 2>&1; echo "racecheck2 rc=$?"
grep -a "RACECHECK SUMMARY\|passed\|failed\|hazard" gpurun_out/racecheck2.log | tail -8
