#!/bin/bash
mkdir -p gpurun_out
for v in "HBP_HALO_DBG=0" "HBP_HALO_DBG=8" "HBP_HALO_DBG=16" "HBP_HALO_DBG=24" "HBP_HALO_DBG=19"; do
  for sh in 0 1 2; do
   echo "== $v shape $sh" >> gpurun_out/r02c_ablate.log
   env $v HBP_MB_BATCH=256 HBP_MB_SHAPES=$sh timeout 120 python tools/conv_microbench.py 2>&1 | grep eng= >> gpurun_out/r02c_ablate.log
   env $v HBP_MB_NORES=1 HBP_MB_BATCH=256 HBP_MB_SHAPES=$sh timeout 120 python tools/conv_microbench.py 2>&1 | grep eng= | sed 's/^/nores /' >> gpurun_out/r02c_ablate.log
  done
done
cat gpurun_out/r02c_ablate.log
