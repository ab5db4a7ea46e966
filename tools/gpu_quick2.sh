#!/bin/bash
# bring-up of the per-tap (persistent grouped) conv cases, then the quick check
TAG=${1:-q}
mkdir -p gpurun_out
timeout 300 python tools/bringup_conv.py --engine 1 --cases 9,10,11,12,13,14 > gpurun_out/bringup_$TAG.log 2>&1; echo "bringup exit $?" >> gpurun_out/bringup_$TAG.log
cat gpurun_out/bringup_$TAG.log
bash tools/gpu_quick.sh $TAG tl
