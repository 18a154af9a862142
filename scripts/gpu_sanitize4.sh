#!/bin/bash
# late round 2: the rewritten tactile code network (mt_tcn_embed: maps, pair lists, MMA kernels, PDL) and the k-d search
# index under compute-sanitizer (racecheck covers the shared-memory aggregation of k_tcn_kmaps / the compaction kernels)
mkdir -p gpurun_out
timeout 1500 compute-sanitizer --tool memcheck --error-exitcode 7 --print-limit 20 python -m pytest tests/test_gpu_tcn.py -m gpu -q -x -k "not oracle and not api" > gpurun_out/memcheck4.log 2>&1; echo "memcheck4 rc=$?"
grep -a "ERROR SUMMARY\|passed\|failed\|Invalid\|out of bounds" gpurun_out/memcheck4.log | tail -6
timeout 1500 compute-sanitizer --tool racecheck --error-exitcode 7 --print-limit 20 python -m pytest tests/test_gpu_tcn.py -m gpu -q -x -k "entry_equals or out_of_range" > gpurun_out/racecheck4.log 2>&1; echo "racecheck4 rc=$?"
grep -a "RACECHECK SUMMARY\|passed\|failed\|hazard" gpurun_out/racecheck4.log | tail -6
timeout 1500 compute-sanitizer --tool memcheck --error-exitcode 7 --print-limit 20 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "heavy_fallback or fused_step_vs_oracle" > gpurun_out/memcheck5.log 2>&1; echo "memcheck5 rc=$?"
grep -a "ERROR SUMMARY\|passed\|failed\|Invalid\|out of bounds" gpurun_out/memcheck5.log | tail -6
