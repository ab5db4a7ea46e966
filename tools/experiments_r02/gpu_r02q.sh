#!/bin/bash
mkdir -p gpurun_out
T=r02q
run() { lbl=$1; shift
  env "$@" timeout 300 python bench.py --steps 10 --warmup 3 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('$lbl hrnet_ms', round(d['roofline']['hrnet_ms'],4), 'parity', d.get('parity_check','')[:10])" | tee -a gpurun_out/${T}_variants.log
}
run v8 HBP_X=0
run v4x2 HBP_HALO_DBG=32
run nostores HBP_HALO_DBG=1
run noepilogue HBP_HALO_DBG=8
for d in 0 32; do
  echo "== HBP_HALO_DBG=$d" | tee -a gpurun_out/${T}_mb.log
  HBP_HALO_DBG=$d HBP_MB_SHAPES=0,1,2,4,5,7 timeout 120 python tools/conv_microbench.py 2>/dev/null | grep "eng=" | tee -a gpurun_out/${T}_mb.log
done
