#!/bin/bash
# One GPU-box session: conv bring-up (both engines), the gpu test tier.  Everything is
# wrapped in its own timeout and logs to gpurun_out/.
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/smi.txt 2>&1
L=gpurun_out/bringup.log
echo "== conv bring-up engine 1 (tcgen05: halo mode for 3x3 s1, per-tap TMA otherwise)" > $L
timeout 300 python tools/bringup_conv.py --engine 1 >> $L 2>&1; echo "exit $?" >> $L
echo "== conv bring-up engine 1, HBP_CONV_HALO=0 (per-tap everywhere)" >> $L
HBP_CONV_HALO=0 timeout 300 python tools/bringup_conv.py --engine 1 --cases 1,2,5,15 >> $L 2>&1; echo "exit $?" >> $L
echo "== conv bring-up engine 0 (SIMT)" >> $L
timeout 300 python tools/bringup_conv.py --engine 0 --cases 0,2,5,9,11,16,17 >> $L 2>&1; echo "exit $?" >> $L
cat $L
echo "== pytest gpu"
timeout 1200 python -m pytest tests -m gpu -q --timeout 600 2>&1 | tail -40 | tee gpurun_out/pytest_gpu.log
