#!/bin/bash
TAG=${1:-r01h}
mkdir -p gpurun_out
L=gpurun_out/bringup_$TAG.log
HBP_CONV_TRACE=1 timeout 300 python tools/bringup_conv.py --engine 1 > $L 2>&1; echo "bringup exit $?" >> $L
grep -v "^\[taps\|^\[trace" $L | cut -c1-250
HBP_MB_SHAPES=0,1,2,3,4,6 timeout 300 python tools/conv_microbench.py > gpurun_out/mb_$TAG.log 2>&1; echo "mb rc=$?"; cat gpurun_out/mb_$TAG.log
HBP_MB_ITERS=4 HBP_MB_SHAPES=0,1,2,3 HBP_CONV_TRACE=1 timeout 300 python tools/conv_microbench.py > gpurun_out/mb_trace_$TAG.log 2>&1; echo "trace rc=$?"
grep "^\[trace\|^\[plan" gpurun_out/mb_trace_$TAG.log | cut -c1-330
