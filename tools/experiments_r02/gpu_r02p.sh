#!/bin/bash
mkdir -p gpurun_out
T=r02p
for P in 64 32; do
for d in 0 1 8 16 24; do
  echo "== P=$P HBP_HALO_DBG=$d" | tee -a gpurun_out/${T}_layer1_ablate.log
  HBP_MB_BATCH=$P HBP_HALO_DBG=$d HBP_MB_SHAPES=4,5,7 timeout 120 python tools/conv_microbench.py 2>/dev/null | grep "eng=" | tee -a gpurun_out/${T}_layer1_ablate.log
done
done
HBP_MB_BATCH=64 HBP_CONV_TRACE=1 HBP_MB_ITERS=0 HBP_MB_SHAPES=5,7 timeout 120 python tools/conv_microbench.py > /dev/null 2> gpurun_out/${T}_trace.log
