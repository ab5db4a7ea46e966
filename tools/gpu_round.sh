#!/bin/bash
# One GPU-box session: conv bring-up, the gpu test tier.  Everything is wrapped in
# its own timeout and logs to gpurun_out/.
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/smi.txt 2>&1
L=gpurun_out/bringup.log
echo "== conv bring-up engine 0" > $L
timeout 300 python tools/bringup_conv.py --engine 0 >> $L 2>&1; echo "exit $?" >> $L
echo "== conv bring-up engine 1, stride-1 cases" >> $L
timeout 300 python tools/bringup_conv.py --engine 1 --cases 0,1,2,3,4,5,6,7,8,9,10,15 >> $L 2>&1; echo "exit $?" >> $L
for mode in 1 2; do
  echo "== conv bring-up engine 1 stride-2 cases, HBP_TMA_STRIDE_MODE=$mode" >> $L
  HBP_TMA_STRIDE_MODE=$mode timeout 200 python tools/bringup_conv.py --engine 1 --cases 11,12,13,14 >> $L 2>&1; echo "exit $?" >> $L
done
cat $L
echo "== pytest gpu"
timeout 1200 python -m pytest tests -m gpu -q --timeout 600 2>&1 | tail -60 | tee gpurun_out/pytest_gpu.log
