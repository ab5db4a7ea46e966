#!/bin/bash
mkdir -p gpurun_out
T=r02o
run() { # label, env...
  lbl=$1; shift
  env "$@" timeout 300 python bench.py --steps 10 --warmup 3 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('$lbl hrnet_ms', round(d['roofline']['hrnet_ms'],4), 'parity', d.get('parity_check','')[:10])" | tee -a gpurun_out/${T}_variants.log
}
run base HBP_L2_HINTS=0 HBP_REVERSE=0
run rev HBP_L2_HINTS=0 HBP_REVERSE=1
run hints HBP_L2_HINTS=1 HBP_REVERSE=0
run rev+hints HBP_L2_HINTS=1 HBP_REVERSE=1
run rev+hints+keep HBP_L2_HINTS=3 HBP_REVERSE=1
timeout 900 python -m pytest tests/test_gpu_hrnet_parity.py tests/test_gpu_parity.py -m gpu -x -q --timeout 600 -k "hrnet or pipeline" 2>&1 | tail -3 | tee gpurun_out/${T}_pytest.log
HBP_TIMELINE=1 timeout 300 python bench.py --steps 3 --warmup 3 > /dev/null 2> gpurun_out/${T}_timeline.log; grep -c "^\[tl\]" gpurun_out/${T}_timeline.log
