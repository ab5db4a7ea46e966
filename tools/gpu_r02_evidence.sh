#!/bin/bash
# Round-2 evidence pass on one GPU (no ncu).  usage: tools/gpu_r02_evidence.sh TAG
TAG=${1:-r02d}
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -q -s --timeout 600 2>&1 | grep -v "^\[plan\|^\[pgroup\|^\[chain" > gpurun_out/${TAG}_pytest_full.log; tail -2 gpurun_out/${TAG}_pytest_full.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -1
timeout 900 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/${TAG}_bench_reference.json 2> gpurun_out/${TAG}_bench.err
cut -c1-200 gpurun_out/${TAG}_bench_reference.json
timeout 900 python bench.py --steps 20 --warmup 3 > gpurun_out/${TAG}_bench.json 2>> gpurun_out/${TAG}_bench.err
echo "bench rc=$?"; cut -c1-300 gpurun_out/${TAG}_bench.json
timeout 600 python bench.py --config 2 --steps 20 --warmup 3 > gpurun_out/${TAG}_bench_c2.json 2>> gpurun_out/${TAG}_bench.err; cut -c1-200 gpurun_out/${TAG}_bench_c2.json
timeout 900 python bench.py --config 3 --steps 5 --warmup 3 > gpurun_out/${TAG}_bench_c3.json 2>> gpurun_out/${TAG}_bench.err; cut -c1-200 gpurun_out/${TAG}_bench_c3.json
grep -c Traceback gpurun_out/${TAG}_bench.err
