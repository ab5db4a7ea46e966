#!/bin/bash
mkdir -p gpurun_out
T=${1:-r02z}
run() { lbl=$1; shift
  env "$@" timeout 300 python bench.py --steps 10 --warmup 3 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('$lbl hrnet_ms', round(d['roofline']['hrnet_ms'],4), 'parity', d.get('parity_check','')[:10])" | tee -a gpurun_out/${T}_variants.log
}
for s in 65,36,47 64,39,45 64,41,43 60,45,43 64,48,36 77,28,43 56,49,43 66,38,44 65,40,43 68,36,44 64,36,48 62,38,48; do
  run s3_$s HBP_BRANCH_SHARE3=$s
done
for s in 56,30,36,26 54,32,36,26 58,30,34,26 56,28,38,26 52,32,38,26 56,32,32,28; do
  run s4_$s HBP_BRANCH_SHARE3=65,36,47 HBP_BRANCH_SHARE4=$s
done
