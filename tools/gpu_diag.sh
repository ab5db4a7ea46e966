#!/bin/bash
# diagnostic pass: phase trace of the conv shape classes, the gpu test tier, launch list of the bench
TAG=${1:-r01f}
mkdir -p gpurun_out
HBP_CONV_TRACE=1 timeout 300 python tools/conv_microbench.py > gpurun_out/mb_trace_$TAG.log 2>&1; echo "trace rc=$?"
timeout 300 python tools/conv_microbench.py > gpurun_out/mb_$TAG.log 2>&1; echo "mb rc=$?"; cat gpurun_out/mb_$TAG.log
timeout 900 python -m pytest tests -m gpu -q --timeout 600 2>&1 | tail -5 | tee gpurun_out/pytest_gpu_$TAG.log
python bench.py --steps 20 --warmup 3 > gpurun_out/bench_$TAG.json 2> gpurun_out/bench_$TAG.err; echo "bench rc=$?"; cat gpurun_out/bench_$TAG.json
python bench.py --steps 2 --warmup 3 > gpurun_out/plain.log 2>&1 &&
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 3000 --csv \
    --log-file gpurun_out/launches_$TAG.csv python bench.py --steps 2 --warmup 3 > gpurun_out/ncu_list.log 2>&1
echo "ncu list rc=$?"
