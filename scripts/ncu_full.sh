#!/bin/bash
# one ncu --set full capture of the step kernels (launch 8..10 of each) -> gpurun_out/prof_*.ncu-rep
mkdir -p gpurun_out
python scripts/variants.py > gpurun_out/variants.json 2> gpurun_out/variants.err; tail -c 2500 gpurun_out/variants.json; tail -3 gpurun_out/variants.err
for k in k_step_a k_step_b k_cosine_rows; do
  timeout 900 ncu --set full --clock-control none --import-source on -k regex:$k -s 8 -c 2 -f -o gpurun_out/prof_$k \
     python bench.py --steps 4 --warmup 8 --no-cpu > gpurun_out/ncu_$k.log 2>&1
  echo "$k rc=$?"
done
ls -la gpurun_out/*.ncu-rep
