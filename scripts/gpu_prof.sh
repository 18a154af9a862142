#!/bin/bash
# ncu --set full of one kernel inside scripts/step_ab.py (KERNEL=regex, SKIP=launches of it to skip, LIB=variants/x.so optional)
mkdir -p gpurun_out
[ -n "${LIB:-}" ] && export MIDAS_B200_LIB=$PWD/$LIB
AB_STEPS=${AB_STEPS:-40} timeout 900 ncu --set full --clock-control none --import-source on -k regex:${KERNEL} -s ${SKIP:-30} -c 1 -f \
   -o gpurun_out/prof_${TAG:-$KERNEL} python scripts/step_ab.py > gpurun_out/ncu_${TAG:-$KERNEL}.log 2>&1
echo "ncu ${KERNEL} rc=$?"; tail -2 gpurun_out/ncu_${TAG:-$KERNEL}.log
