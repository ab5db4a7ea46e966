#!/bin/bash
mkdir -p gpurun_out
T=${1:-r02f}
timeout 900 python -m pytest tests/test_gpu_hrnet_parity.py tests/test_gpu_parity.py -m gpu -x -q --timeout 600 -k "hrnet or pipeline" 2>&1 | tail -25 | tee gpurun_out/${T}_pytest_hrnet.log
HBP_CONV_TRACE=1 timeout 300 python bench.py --steps 3 --warmup 3 2>&1 >/dev/null | grep "\[chain\]" | sort | uniq | head -20 | tee gpurun_out/${T}_chain_plans.log
timeout 600 python bench.py --steps 10 --warmup 3 > gpurun_out/${T}_bench.json 2> gpurun_out/${T}_bench.err; echo "bench rc=$?"
python - <<PY
import json
d = json.load(open("gpurun_out/${T}_bench.json"))
print("value", d["value"], "e2e", d["e2e"]["value"], "hrnet_ms", d["roofline"]["hrnet_ms"], "frac", d["roofline"]["frac"])
PY
HBP_CHAIN=0 timeout 600 python bench.py --steps 10 --warmup 3 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('nochain hrnet_ms', d['roofline']['hrnet_ms'])"
HBP_TIMELINE=1 timeout 300 python bench.py --steps 3 --warmup 2 > /dev/null 2> gpurun_out/${T}_timeline.log; grep -c "^\[tl\]" gpurun_out/${T}_timeline.log
