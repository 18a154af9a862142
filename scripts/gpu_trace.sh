#!/bin/bash
# gpurun driver: scripts/step_trace.py (graph form, then stream form) for every variants/*.so built with -DMT_TRACE=1
mkdir -p gpurun_out
: > gpurun_out/trace.jsonl
for lib in variants/*.so; do
  MIDAS_B200_LIB=$PWD/$lib timeout 300 python scripts/step_trace.py 2> gpurun_out/trace.err | tee -a gpurun_out/trace.jsonl || tail -5 gpurun_out/trace.err
  if [ -n "${STREAM:-}" ]; then
  AB_STREAM=1 MIDAS_B200_LIB=$PWD/$lib timeout 300 python scripts/step_trace.py 2> gpurun_out/trace.err | tee -a gpurun_out/trace.jsonl || tail -5 gpurun_out/trace.err
  fi
done
