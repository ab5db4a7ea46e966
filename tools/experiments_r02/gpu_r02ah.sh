#!/bin/bash
mkdir -p gpurun_out
T=r02ah
for sms in 47 36 30; do
  echo "== SMS=$sms" | tee -a gpurun_out/${T}_b2.log
  HBP_HALO_SMS=$sms HBP_MB_SHAPES=2 HBP_CONV_TRACE=1 timeout 120 python tools/conv_microbench.py 2>&1 | grep "eng=\|\[plan\]" | cut -c1-200 | tee -a gpurun_out/${T}_b2.log
done
timeout 900 python -m pytest tests/test_gpu_hrnet_parity.py tests/test_gpu_parity.py -m gpu -x -q --timeout 600 -k "hrnet or pipeline or conv" 2>&1 | tail -3 | tee gpurun_out/${T}_pytest.log
run() { lbl=$1; shift
  env "$@" timeout 300 python bench.py --steps 10 --warmup 3 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('$lbl hrnet_ms', round(d['roofline']['hrnet_ms'],4), 'parity', d.get('parity_check','')[:10])" | tee -a gpurun_out/${T}_variants.log
}
run base HBP_X=0
run s3_46_26_28 HBP_BRANCH_SHARE3=0.46,0.26,0.28
run s3_46_28_26 HBP_BRANCH_SHARE3=0.46,0.28,0.26
run s3_48_26_26 HBP_BRANCH_SHARE3=0.48,0.26,0.26
run s3_44_28_28 HBP_BRANCH_SHARE3=0.44,0.28,0.28
run s4_40_22_20_18 HBP_BRANCH_SHARE4=0.40,0.22,0.20,0.18
run s4_40_20_20_20 HBP_BRANCH_SHARE4=0.40,0.20,0.20,0.20
