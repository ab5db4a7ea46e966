#!/bin/bash
# gpu test tier + bench, short
TAG=${1:-r01}
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q --timeout 600 2>&1 | tail -5 | tee gpurun_out/pytest_gpu_$TAG.log
python bench.py --steps 20 --warmup 3 > gpurun_out/bench_$TAG.json 2> gpurun_out/bench_$TAG.err; echo "bench rc=$?"
python - <<PY
import json
d=json.load(open("gpurun_out/bench_$TAG.json"))
print({k:d[k] for k in ("value","ms_per_step")}, d["e2e"]["value"], d["roofline"]["frac"], d["roofline"]["hrnet_ms"], d["stages_ms"])
PY
tail -3 gpurun_out/bench_$TAG.err
