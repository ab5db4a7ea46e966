#!/bin/bash
# ncu --set full (dense PC sampling) of the halo conv kernel on the biggest shape class
TAG=${1:-r01v}
mkdir -p gpurun_out
HBP_MB_ITERS=4 HBP_MB_SHAPES=0 timeout 300 python tools/conv_microbench.py > gpurun_out/plain.log 2>&1 &&
HBP_MB_ITERS=4 HBP_MB_SHAPES=0 timeout 600 ncu --set full --warp-sampling-interval 0 --clock-control none --import-source on -k regex:conv_umma_halo -s 2 -c 2 \
    -o gpurun_out/prof_halo_$TAG -f python tools/conv_microbench.py > gpurun_out/ncu_full.log 2>&1
echo "ncu rc=$?"; tail -3 gpurun_out/ncu_full.log
