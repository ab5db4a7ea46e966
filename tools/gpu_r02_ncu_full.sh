#!/bin/bash
# ncu --set full of representative conv kernels of one eager HRNet forward (64 crops); reports kept under 64 MiB in total.
# usage: tools/gpu_r02_ncu_full.sh TAG
TAG=${1:-r02c}
mkdir -p gpurun_out
HBP_NO_GRAPH=1 timeout 120 python tools/hrnet_only.py 2 || exit 1
# second forward: halo launches 3.. = layer1 (1x1 / 3x3 / 1x1+residual), 36.. = stage 3 (32-, 64-, 128-channel branch convs)
HBP_NO_GRAPH=1 timeout 600 ncu --set full --clock-control none -k regex:conv_umma_halo -s 260 -c 6 \
    -o gpurun_out/prof_halo_layer1_$TAG -f python tools/hrnet_only.py 2 > gpurun_out/ncu_full_$TAG.log 2>&1
echo "ncu layer1 rc=$?"
HBP_NO_GRAPH=1 timeout 600 ncu --set full --clock-control none -k regex:conv_umma_halo -s 293 -c 12 \
    -o gpurun_out/prof_halo_stage3_$TAG -f python tools/hrnet_only.py 2 >> gpurun_out/ncu_full_$TAG.log 2>&1
echo "ncu stage3 rc=$?"
HBP_NO_GRAPH=1 timeout 600 ncu --set full --clock-control none -k regex:conv_umma_pgroup -s 44 -c 4 \
    -o gpurun_out/prof_pgroup_$TAG -f python tools/hrnet_only.py 2 >> gpurun_out/ncu_full_$TAG.log 2>&1
echo "ncu pgroup rc=$?"
du -sh gpurun_out
