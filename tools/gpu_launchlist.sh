#!/bin/bash
TAG=${1:-r01x}
mkdir -p gpurun_out
python bench.py --steps 2 --warmup 3 > gpurun_out/plain.log 2>&1 &&
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 3000 --csv \
    --log-file gpurun_out/launches_$TAG.csv python bench.py --steps 2 --warmup 3 > gpurun_out/ncu_list.log 2>&1
echo "ncu list rc=$?"
