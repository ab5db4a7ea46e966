#!/bin/bash
mkdir -p gpurun_out
T=r02ag
for kb in 216 226; do
  echo "== SMEM_KB=$kb SMS=47" | tee -a gpurun_out/${T}_b2.log
  HBP_HALO_SMEM_KB=$kb HBP_HALO_SMS=47 HBP_MB_SHAPES=2 HBP_CONV_TRACE=1 timeout 120 python tools/conv_microbench.py 2>&1 | grep "eng=\|\[plan\]" | cut -c1-200 | tee -a gpurun_out/${T}_b2.log
done
run() { lbl=$1; shift
  env "$@" timeout 300 python bench.py --steps 10 --warmup 3 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('$lbl hrnet_ms', round(d['roofline']['hrnet_ms'],4), 'parity', d.get('parity_check','')[:10])" | tee -a gpurun_out/${T}_variants.log
}
run kb216 HBP_HALO_SMEM_KB=216
run kb226 HBP_HALO_SMEM_KB=226
run kb226_s3_44_26_30 HBP_HALO_SMEM_KB=226 HBP_BRANCH_SHARE3=0.44,0.26,0.30
run kb226_s3_46_26_28 HBP_HALO_SMEM_KB=226 HBP_BRANCH_SHARE3=0.46,0.26,0.28
run kb226_nores HBP_HALO_SMEM_KB=226 HBP_HALO_RES_SMEM=1
