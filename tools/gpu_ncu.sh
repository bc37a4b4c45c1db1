#!/bin/bash
# Usage: tools/gpu_ncu.sh TAG KERNEL_REGEX [profile_step args...]   -- one `ncu --set full` capture (3 launches max)
TAG=$1; RX=$2; shift 2
mkdir -p gpurun_out
python tools/profile_step.py "$@" > gpurun_out/plain.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:$RX -c 2 -f -o gpurun_out/prof_$TAG \
    python tools/profile_step.py "$@" > gpurun_out/ncu_$TAG.log 2>&1
echo "ncu rc $?"; tail -3 gpurun_out/ncu_$TAG.log
