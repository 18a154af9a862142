#!/bin/bash
# gpurun --gpus N -- 'bash scripts/gpu_multi_bench.sh N' : N-GPU bench line only
N=${1:-2}
mkdir -p gpurun_out
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29542 \
  bench.py --gpus $N --steps ${STEPS:-50} --warmup 5 --no-cpu > gpurun_out/bench_${N}gpu.json 2> gpurun_out/bench_${N}gpu.err
echo "bench rc=$?"
python - $N <<'PY'
import json,sys
d=json.load(open('gpurun_out/bench_%sgpu.json'%sys.argv[1]))
print(d['value'], d['ms_per_step'], 'e2e', d['e2e']['value'])
for r in d['per_rank']: print({k.split(' ')[0]: round(v,4) for k,v in r.items()})
print(d['filter']['step_ms_every_5th'])
PY
tail -3 gpurun_out/bench_${N}gpu.err
