#!/bin/bash
# Evidence pass on one GPU: gpu test tier, bench (both arms), ncu launch list (time + DRAM bytes) of one graph
# replay, ncu --set full captures of the dominant kernels.  usage: tools/gpu_final.sh TAG
TAG=${1:-r01}
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q --timeout 600 2>&1 | tail -5 | tee gpurun_out/pytest_$TAG.log
timeout 900 python bench.py --steps 20 --warmup 3 > gpurun_out/bench_$TAG.json 2> gpurun_out/bench_$TAG.err
echo "bench rc=$?"; cut -c1-600 gpurun_out/bench_$TAG.json
timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_ref_$TAG.json 2>> gpurun_out/bench_$TAG.err
cut -c1-400 gpurun_out/bench_ref_$TAG.json
# launch list of the same command (cold-cache, serialised: shares, not absolutes); the last graph replay = one forward
timeout 900 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -c 4000 --csv \
    --log-file gpurun_out/launches_$TAG.csv python bench.py --steps 2 --warmup 3 > gpurun_out/ncu_list.log 2>&1
echo "ncu list rc=$?"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:conv_umma_halo -s 60 -c 4 \
    -o gpurun_out/prof_halo_$TAG -f python bench.py --steps 2 --warmup 3 > gpurun_out/ncu_full.log 2>&1
echo "ncu halo rc=$?"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:conv_umma_pgroup -s 8 -c 3 \
    -o gpurun_out/prof_pgroup_$TAG -f python bench.py --steps 2 --warmup 3 >> gpurun_out/ncu_full.log 2>&1
echo "ncu pgroup rc=$?"
timeout 300 python tools/w48_time.py > gpurun_out/w48_time_$TAG.log 2>&1; cat gpurun_out/w48_time_$TAG.log
ls gpurun_out | wc -l
