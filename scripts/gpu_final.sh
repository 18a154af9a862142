#!/bin/bash
# final visit of a round: smoke, default bench line, ncu launch list of the same command
mkdir -p gpurun_out
timeout 300 python __graft_entry__.py smoke > gpurun_out/smoke.log 2>&1; echo "smoke rc=$?" >> gpurun_out/smoke.log; tail -2 gpurun_out/smoke.log
timeout 900 python bench.py > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "bench rc=$?"; tail -3 gpurun_out/bench.err
python - <<'PY'
import json
d=json.load(open('gpurun_out/bench.json'))
print(d['value'], d['ms_per_step'], 'e2e', d['e2e']['value'], d['roofline']['frac'], d['clocks'], d['roofline']['sweep']['kernels_ms'])
PY
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/launches.csv \
    python bench.py --steps 30 --warmup 5 --no-cpu > gpurun_out/bench_ncu.log 2>&1; echo "ncu launches rc=$?"
