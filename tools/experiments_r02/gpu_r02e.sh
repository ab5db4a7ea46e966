#!/bin/bash
mkdir -p gpurun_out
HBP_CONV_TRACE=1 HBP_MB_BATCH=256 HBP_MB_SHAPES=0,1 HBP_MB_ITERS=0 timeout 300 python tools/conv_microbench.py > gpurun_out/r02e_trace_P256.log 2>&1
HBP_CONV_TRACE=1 HBP_MB_NORES=1 HBP_MB_BATCH=256 HBP_MB_SHAPES=0 HBP_MB_ITERS=0 timeout 300 python tools/conv_microbench.py > gpurun_out/r02e_trace_nores_P256.log 2>&1
for P in 64 256; do
  HBP_MB_BATCH=$P HBP_MB_SHAPES=0,1,2,3 timeout 200 python tools/conv_microbench.py 2>&1 | grep eng= | sed "s/^/P=$P /" | tee -a gpurun_out/r02e_mb.log
done
timeout 1200 python -m pytest tests -m gpu -q --timeout 900 2>&1 | tail -25 | tee gpurun_out/r02e_pytest.log
