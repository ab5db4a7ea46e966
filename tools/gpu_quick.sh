#!/bin/bash
# quick GPU check: gpu test tier + short bench (+ optional timeline).  usage: tools/gpu_quick.sh tag [timeline]
TAG=${1:-q}
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -x --timeout 600 2>&1 | tail -15 | tee gpurun_out/pytest_$TAG.log
timeout 600 python bench.py --steps 10 --warmup 3 > gpurun_out/bench_$TAG.json 2> gpurun_out/bench_$TAG.err
echo "bench rc=$?"; python - <<PY
import json
try:
    d = json.load(open("gpurun_out/bench_$TAG.json"))
    print("value", d["value"], "e2e", d["e2e"]["value"], "hrnet_ms", d["roofline"]["hrnet_ms"], "frac", d["roofline"]["frac"])
    print(d.get("stages_ms"))
    print({k: (round(v["frac"], 3), round(v["ms"], 3)) for k, v in d.get("stage_rooflines", {}).items()})
except Exception as e:
    print("no bench json", e)
PY
tail -3 gpurun_out/bench_$TAG.err
if [ -n "$2" ]; then
  HBP_TIMELINE=1 timeout 300 python bench.py --steps 3 --warmup 2 > /dev/null 2> gpurun_out/timeline_$TAG.log
  grep -c "^\[tl\]" gpurun_out/timeline_$TAG.log
fi
