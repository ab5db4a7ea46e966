#!/bin/bash
# r02a: baseline sanity + per-tile slope of the branch conv kernels (time vs batch) + per-CTA phase trace
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv > gpurun_out/r02a_smi.log
for P in 64 128 256; do
  HBP_MB_BATCH=$P HBP_MB_SHAPES=0,1,2,3 timeout 300 python tools/conv_microbench.py > gpurun_out/r02a_mb_P$P.log 2>&1
  HBP_MB_NORES=1 HBP_MB_BATCH=$P HBP_MB_SHAPES=0,1,2,3 timeout 300 python tools/conv_microbench.py > gpurun_out/r02a_mb_nores_P$P.log 2>&1
done
HBP_CONV_TRACE=1 HBP_MB_BATCH=64 HBP_MB_SHAPES=0,1,2,3 HBP_MB_ITERS=0 timeout 300 python tools/conv_microbench.py > gpurun_out/r02a_trace.log 2>&1
timeout 600 python bench.py --steps 10 --warmup 3 > gpurun_out/r02a_bench.json 2> gpurun_out/r02a_bench.err
echo bench rc=$?
tail -2 gpurun_out/r02a_mb_P64.log
