#!/bin/bash
# Usage: tools/gpu_ncu2.sh TAG KERNEL_REGEX SKIP COUNT [profile_step args...]
TAG=$1; RX=$2; SKIP=$3; CNT=$4; shift 4
mkdir -p gpurun_out
python tools/profile_step.py "$@" > gpurun_out/plain.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k "regex:$RX" -s $SKIP -c $CNT -f -o gpurun_out/prof_$TAG \
    python tools/profile_step.py "$@" > gpurun_out/ncu_$TAG.log 2>&1
echo "ncu rc $?"; tail -3 gpurun_out/ncu_$TAG.log
