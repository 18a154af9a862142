#!/bin/bash
# full GPU visit (1 GPU): build check, tests, smoke, bench (+ reference arm), launch list, ncu --set full of every step kernel,
# the batched query (tensor pipe), sanitizer over a test subset.  Outputs under gpurun_out/.
set -u
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.build()" > gpurun_out/build.log 2>&1
timeout 1200 python -m pytest tests -m gpu -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
tail -4 gpurun_out/pytest_gpu.log
timeout 300 python __graft_entry__.py smoke > gpurun_out/smoke.log 2>&1; echo "smoke rc=$?" >> gpurun_out/smoke.log; tail -2 gpurun_out/smoke.log
nvidia-smi --query-gpu=index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap --format=csv -lms 200 > gpurun_out/clocks.csv &
SMI=$!
timeout 900 python bench.py --steps 100 --warmup 5 > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "bench rc=$?"; python scripts/show_bench.py gpurun_out/bench.json | head -12
kill $SMI
timeout 600 python bench.py --impl reference --steps 20 --warmup 5 > gpurun_out/bench_reference.json 2>> gpurun_out/bench.err; echo "reference rc=$?"; tail -c 600 gpurun_out/bench_reference.json
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 800 --csv --log-file gpurun_out/launches.csv \
    python bench.py --steps 30 --warmup 5 --no-cpu --no-converged > gpurun_out/bench_ncu.log 2>&1; echo "ncu launches rc=$?"
for k in k_step_a k_step_meshq k_step_meshq2 k_step_nnq k_step_bw k_codebook_query; do
  AB_STEPS=40 timeout 900 ncu --set full --clock-control none --import-source on -k regex:"$k\\b" -s 30 -c 1 -f -o gpurun_out/prof_$k \
     python scripts/step_ab.py > gpurun_out/ncu_$k.log 2>&1
  echo "$k rc=$?"
done
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_codebook_gemm_tma -s 2 -c 1 -f -o gpurun_out/prof_k_codebook_gemm_tma \
     python scripts/gemm_prof.py > gpurun_out/ncu_gemm.log 2>&1; echo "gemm rc=$?"
python scripts/gemm_prof.py > gpurun_out/gemm.log 2>&1; tail -2 gpurun_out/gemm.log
python scripts/pin_ab.py > gpurun_out/pin.json 2> gpurun_out/pin.err; tail -1 gpurun_out/pin.json
timeout 900 compute-sanitizer --tool memcheck python -m pytest tests/test_gpu_parity.py tests/test_gpu_reference_loop.py -m gpu -q -x -k "fused_step_vs_oracle or dbscan or prune_vs or batched_tensor_core or cluster_centers or reference_loop" > gpurun_out/memcheck.log 2>&1; tail -3 gpurun_out/memcheck.log
