#!/bin/bash
# Round-2 evidence pass on one GPU.  usage: tools/gpu_r02_evidence.sh TAG
TAG=${1:-r02a}
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -q -s --timeout 600 2>&1 | grep -v "^\[plan\|^\[pgroup\|^\[chain" > gpurun_out/${TAG}_pytest_full.log; tail -3 gpurun_out/${TAG}_pytest_full.log
timeout 900 python bench.py --steps 20 --warmup 3 > gpurun_out/${TAG}_bench.json 2> gpurun_out/${TAG}_bench.err
echo "bench rc=$?"; cut -c1-400 gpurun_out/${TAG}_bench.json
timeout 900 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/${TAG}_bench_reference.json 2>> gpurun_out/${TAG}_bench.err
cut -c1-300 gpurun_out/${TAG}_bench_reference.json
timeout 600 python bench.py --config 2 --steps 20 --warmup 3 > gpurun_out/${TAG}_bench_c2.json 2>> gpurun_out/${TAG}_bench.err; cut -c1-300 gpurun_out/${TAG}_bench_c2.json
timeout 900 python bench.py --config 3 --steps 5 --warmup 3 > gpurun_out/${TAG}_bench_c3.json 2>> gpurun_out/${TAG}_bench.err; cut -c1-300 gpurun_out/${TAG}_bench_c3.json
timeout 900 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -c 8000 --csv \
    --log-file gpurun_out/${TAG}_launches.csv python bench.py --steps 2 --warmup 3 > gpurun_out/${TAG}_ncu_list.log 2>&1
echo "ncu list rc=$?"; grep -c "gpu__time_duration" gpurun_out/${TAG}_launches.csv
