#!/bin/bash
# branch SM shares: whole-network bench for a few share vectors + microbench of each branch class at its SM budget
TAG=${1:-r01t}
mkdir -p gpurun_out
for sh in "0.34,0.21,0.19,0.26" "0.40,0.20,0.17,0.23" "0.30,0.22,0.20,0.28" "1,1,1,1"; do
  echo "== share $sh"
  HBP_BRANCH_SHARE=$sh python bench.py --steps 10 --warmup 3 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('hrnet_ms', d['roofline']['hrnet_ms'], 'value', d['value'])"
done
for cfg in "0 50" "1 31" "2 28" "3 38"; do
  set -- $cfg
  HBP_HALO_SMS=$2 HBP_MB_ITERS=20 HBP_MB_SHAPES=$1 HBP_CONV_TRACE=1 timeout 120 python tools/conv_microbench.py 2>&1 | grep "plan\|TFLOP/s " | cut -c1-250
done
