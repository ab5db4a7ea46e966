#!/bin/bash
mkdir -p gpurun_out
T=${1:-r02g}
timeout 900 python -m pytest tests/test_gpu_hrnet_parity.py tests/test_gpu_parity.py -m gpu -x -q --timeout 600 -k "hrnet or pipeline" 2>&1 | tail -5 | tee gpurun_out/${T}_pytest_hrnet.log
for v in 0 3 7; do
HBP_CHAIN_DBG=$v timeout 300 python bench.py --steps 10 --warmup 3 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('chain_dbg=$v hrnet_ms', d['roofline']['hrnet_ms'])" | tee -a gpurun_out/${T}_variants.log
done
HBP_CHAIN=0 timeout 300 python bench.py --steps 10 --warmup 3 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('nochain hrnet_ms', d['roofline']['hrnet_ms'])" | tee -a gpurun_out/${T}_variants.log
HBP_TIMELINE=1 timeout 300 python bench.py --steps 3 --warmup 2 > /dev/null 2> gpurun_out/${T}_timeline.log; grep -c "^\[tl\]" gpurun_out/${T}_timeline.log
