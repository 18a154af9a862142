#!/bin/bash
# gpurun driver: [TESTS=1: pytest -m gpu with the default library], then scripts/step_ab.py for every variants/*.so
mkdir -p gpurun_out
if [ "${TESTS:-0}" = "1" ]; then
timeout 1200 python -m pytest tests -m gpu -q -x > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
tail -5 gpurun_out/pytest_gpu.log
fi
: > gpurun_out/ab.jsonl
for lib in variants/*.so; do
  MIDAS_B200_LIB=$PWD/$lib timeout 300 python scripts/step_ab.py 2> gpurun_out/ab.err | tee -a gpurun_out/ab.jsonl || tail -5 gpurun_out/ab.err
done
if [ -n "${NOFLUSH:-}" ]; then
for lib in variants/*.so; do
  AB_NOFLUSH=1 MIDAS_B200_LIB=$PWD/$lib timeout 300 python scripts/step_ab.py 2> gpurun_out/ab.err | tee -a gpurun_out/ab.jsonl
done
fi
