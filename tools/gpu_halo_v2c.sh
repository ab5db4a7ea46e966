#!/bin/bash
TAG=${1:-r01j}
mkdir -p gpurun_out
L=gpurun_out/bringup_$TAG.log
HBP_CONV_TRACE=1 timeout 300 python tools/bringup_conv.py --engine 1 --cases 1,2,3,4,5,8,15,16,17,18,19,20 > $L 2>&1; echo "bringup exit $?" >> $L
grep -v "^\[taps\|^\[trace\|^\[steady" $L | cut -c1-250
for cfg in "A=1" "HBP_HALO_M=2" "HBP_HALO_M=2 HBP_HALO_S=1" "HBP_HALO_S=1"; do
  echo "=== cfg: $cfg"
  env $cfg HBP_MB_ITERS=20 HBP_MB_SHAPES=0,1,2,3 HBP_CONV_TRACE=1 timeout 300 python tools/conv_microbench.py 2>&1 | grep -v "^\[taps" | cut -c1-420
done > gpurun_out/mb_trace_$TAG.log 2>&1
cat gpurun_out/mb_trace_$TAG.log
