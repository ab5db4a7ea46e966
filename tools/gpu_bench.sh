#!/bin/bash
# bench + ncu evidence on one GPU.  usage: tools/gpu_bench.sh [tag]
TAG=${1:-r01}
mkdir -p gpurun_out
set -o pipefail
python bench.py --steps 20 --warmup 3 > gpurun_out/bench_$TAG.json 2> gpurun_out/bench_$TAG.err
rc=$?
echo "bench rc=$rc"; cat gpurun_out/bench_$TAG.json; tail -5 gpurun_out/bench_$TAG.err
if [ $rc -ne 0 ]; then exit $rc; fi
python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_ref_$TAG.json 2>> gpurun_out/bench_$TAG.err
cat gpurun_out/bench_ref_$TAG.json
# launch list (cold-cache, serialised): shares, not absolutes
python bench.py --steps 2 --warmup 3 > gpurun_out/plain.log 2>&1 &&
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 3000 --csv \
    --log-file gpurun_out/launches_$TAG.csv python bench.py --steps 2 --warmup 3 > gpurun_out/ncu_list.log 2>&1
echo "ncu list rc=$?"
# full capture of the top kernel (3 launches of the tcgen05 conv: the 32->32 3x3 class)
timeout 600 ncu --set full --clock-control none --import-source on -k regex:conv_umma -s 40 -c 3 \
    -o gpurun_out/prof_conv_umma_$TAG -f python bench.py --steps 2 --warmup 3 > gpurun_out/ncu_full.log 2>&1
echo "ncu full rc=$?"
ls -la gpurun_out/
