#!/bin/bash
# compute-sanitizer memcheck / racecheck over a subset of the GPU parity tests (small sizes)
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.build()" > gpurun_out/build.log 2>&1
SEL='aos_soa or nn_far or neighbour_graph or motion_vs or similarity_vs or cosine_ragged or low_var_vs_reference or low_var_heavy or rmse_vs or prune_vs or annealing_vs or cluster_centers or with_prune or all_drifted or query_batched'
timeout 1500 compute-sanitizer --tool memcheck --error-exitcode 7 --print-limit 20 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "$SEL" > gpurun_out/memcheck.log 2>&1; echo "memcheck rc=$?"
grep -a "ERROR SUMMARY\|passed\|failed\|Invalid\|out of bounds" gpurun_out/memcheck.log | tail -8
timeout 900 compute-sanitizer --tool memcheck --error-exitcode 7 --print-limit 20 python -m pytest tests/test_gpu_tcn.py -m gpu -q -x -k "level_counts or cpu_input" > gpurun_out/memcheck_tcn.log 2>&1; echo "memcheck tcn rc=$?"
grep -a "ERROR SUMMARY\|passed\|failed\|Invalid" gpurun_out/memcheck_tcn.log | tail -5
timeout 900 compute-sanitizer --tool racecheck --error-exitcode 7 --print-limit 20 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "with_prune or low_var_heavy or cluster_centers" > gpurun_out/racecheck.log 2>&1; echo "racecheck rc=$?"
grep -a "RACECHECK SUMMARY\|passed\|failed\|hazard" gpurun_out/racecheck.log | tail -8
