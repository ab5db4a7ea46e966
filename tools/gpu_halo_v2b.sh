#!/bin/bash
TAG=${1:-r01i}
mkdir -p gpurun_out
for cfg in "" "HBP_HALO_M=2" "HBP_HALO_ASTAGES=1" ; do
  echo "=== cfg: $cfg"
  env $cfg HBP_MB_ITERS=20 HBP_MB_SHAPES=0,1 HBP_CONV_TRACE=1 timeout 300 python tools/conv_microbench.py 2>&1 | grep -v "^\[taps" | cut -c1-420
done > gpurun_out/mb_trace_$TAG.log 2>&1
cat gpurun_out/mb_trace_$TAG.log
