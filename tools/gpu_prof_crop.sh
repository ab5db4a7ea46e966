#!/bin/bash
mkdir -p gpurun_out
python bench.py --steps 2 --warmup 3 > gpurun_out/plain.log 2>&1 &&
timeout 600 ncu --set full --clock-control none --import-source on -k regex:crop_warp -s 2 -c 2 -o gpurun_out/prof_crop_r01 -f python bench.py --steps 2 --warmup 3 > gpurun_out/ncu_full.log 2>&1
echo "ncu rc=$?"; tail -2 gpurun_out/ncu_full.log
