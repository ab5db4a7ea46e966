#!/bin/bash
mkdir -p gpurun_out
T=r02r
run() { lbl=$1; shift
  env "$@" timeout 300 python bench.py --steps 10 --warmup 3 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('$lbl hrnet_ms', round(d['roofline']['hrnet_ms'],4), 'parity', d.get('parity_check','')[:10])" | tee -a gpurun_out/${T}_variants.log
}
run base HBP_X=0
run smem208 HBP_HALO_SMEM_KB=208
run smem216 HBP_HALO_SMEM_KB=216
for kb in 200 216; do
  echo "== HBP_HALO_SMEM_KB=$kb" | tee -a gpurun_out/${T}_mb.log
  HBP_HALO_SMEM_KB=$kb HBP_MB_SHAPES=0,1,2,4,5,7 timeout 120 python tools/conv_microbench.py 2>/dev/null | grep "eng=" | tee -a gpurun_out/${T}_mb.log
done
HBP_HALO_SMEM_KB=216 HBP_CONV_TRACE=1 timeout 120 python bench.py --steps 1 --warmup 1 2>&1 >/dev/null | grep "\[plan\]" | sort -u > gpurun_out/${T}_plans216.log
