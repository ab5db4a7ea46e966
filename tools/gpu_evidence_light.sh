#!/bin/bash
# bench (both arms) + ncu launch list (time + DRAM bytes) of the same command; no --set full captures.  usage: TAG
TAG=${1:-r01}
mkdir -p gpurun_out
timeout 900 python bench.py --steps 20 --warmup 3 > gpurun_out/bench_$TAG.json 2> gpurun_out/bench_$TAG.err
echo "bench rc=$?"; cut -c1-300 gpurun_out/bench_$TAG.json
timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_ref_$TAG.json 2>> gpurun_out/bench_$TAG.err
cut -c1-200 gpurun_out/bench_ref_$TAG.json
timeout 900 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -c 6000 --csv \
    --log-file gpurun_out/launches_$TAG.csv python bench.py --steps 2 --warmup 3 > gpurun_out/ncu_list.log 2>&1
echo "ncu list rc=$?"; grep -c "gpu__time_duration" gpurun_out/launches_$TAG.csv
