#!/bin/bash
# gpurun driver: cotter-pin stand-in (BASELINE config 5) with forced 64-entry lists and with the upload's own choice,
# then the drill workload (scripts/step_ab.py) for every variants/*.so
mkdir -p gpurun_out
: > gpurun_out/pin.json
for lib in variants/*.so; do
  MIDAS_B200_NBR_K=64 MIDAS_B200_LIB=$PWD/$lib timeout 600 python scripts/pin_ab.py 2> gpurun_out/pin.err | tee -a gpurun_out/pin.json || tail -3 gpurun_out/pin.err
  MIDAS_B200_LIB=$PWD/$lib timeout 600 python scripts/pin_ab.py 2> gpurun_out/pin.err | tee -a gpurun_out/pin.json || tail -3 gpurun_out/pin.err
  AB_GRAPH=1 MIDAS_B200_LIB=$PWD/$lib timeout 600 python scripts/pin_ab.py 2> gpurun_out/pin.err | tee -a gpurun_out/pin.json || tail -3 gpurun_out/pin.err
done
bash scripts/gpu_ab.sh
