#!/bin/bash
mkdir -p gpurun_out
for P in 64 256; do
HBP_CONV_TRACE=1 HBP_MB_BATCH=$P HBP_MB_SHAPES=0,1,2,3 HBP_MB_ITERS=0 timeout 300 python tools/conv_microbench.py > gpurun_out/r02b_trace_P$P.log 2>&1
done
# ablations on the 32-channel branch conv at P=256 (steady state): stores off, bias off, one issuer, fewer stages
for v in "HBP_HALO_DBG=0" "HBP_HALO_DBG=1" "HBP_HALO_DBG=3" "HBP_HALO_ISSUERS=1" "HBP_HALO_RES_SMEM=0" "HBP_HALO_M=2" "HBP_HALO_BUFS=2"; do
  for sh in 0 1 2 3; do
   echo "== $v shape $sh" >> gpurun_out/r02b_ablate.log
   env $v HBP_MB_BATCH=256 HBP_MB_SHAPES=$sh timeout 120 python tools/conv_microbench.py 2>&1 | grep eng= >> gpurun_out/r02b_ablate.log
   env $v HBP_MB_NORES=1 HBP_MB_BATCH=256 HBP_MB_SHAPES=$sh timeout 120 python tools/conv_microbench.py 2>&1 | grep eng= | sed 's/^/nores /' >> gpurun_out/r02b_ablate.log
  done
done
tail -5 gpurun_out/r02b_ablate.log
