#!/bin/bash
# gpurun --gpus N -- 'bash scripts/gpu_multi.sh N' : sharded engine check + N-GPU bench line
N=${1:-2}
mkdir -p gpurun_out
timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29541 \
  scripts/multigpu_check.py > gpurun_out/multigpu_check.log 2>&1
echo "multigpu_check rc=$?" >> gpurun_out/multigpu_check.log
tail -15 gpurun_out/multigpu_check.log
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29542 \
  bench.py --gpus $N --steps 50 --warmup 5 --no-cpu > gpurun_out/bench_${N}gpu.json 2> gpurun_out/bench_${N}gpu.err
echo "bench rc=$?"
cat gpurun_out/bench_${N}gpu.json
tail -5 gpurun_out/bench_${N}gpu.err
