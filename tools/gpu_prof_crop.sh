#!/bin/bash
mkdir -p gpurun_out
python tools/crop_prof.py > gpurun_out/crop_prof.log 2>&1 && cat gpurun_out/crop_prof.log &&
REPS=2 timeout 600 ncu --set full --clock-control none --import-source on -k regex:crop_warp -c 1 -o gpurun_out/prof_crop_${1:-r01b} -f python tools/crop_prof.py > gpurun_out/ncu_crop.log 2>&1
echo "ncu rc=$?"; tail -2 gpurun_out/ncu_crop.log
