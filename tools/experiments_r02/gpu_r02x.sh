#!/bin/bash
mkdir -p gpurun_out
T=${1:-r02x}
run() { lbl=$1; shift
  env "$@" timeout 300 python bench.py --steps 10 --warmup 3 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('$lbl hrnet_ms', round(d['roofline']['hrnet_ms'],4), 'parity', d.get('parity_check','')[:10])" | tee -a gpurun_out/${T}_variants.log
}
run old_model HBP_HALO_WSTREAM_PCT=100
run new_model HBP_X=0
run s3_44_30_26 HBP_BRANCH_SHARE3=0.44,0.30,0.26
run s3_46_32_22 HBP_BRANCH_SHARE3=0.46,0.32,0.22
run s3_48_34_18 HBP_BRANCH_SHARE3=0.48,0.34,0.18
run s3_50_30_20 HBP_BRANCH_SHARE3=0.50,0.30,0.20
run s4_38_26_18_18 HBP_BRANCH_SHARE4=0.38,0.26,0.18,0.18
run s4_40_26_16_18 HBP_BRANCH_SHARE4=0.40,0.26,0.16,0.18
run s4_38_24_14_24 HBP_BRANCH_SHARE4=0.38,0.24,0.14,0.24
HBP_CONV_TRACE=1 timeout 120 python bench.py --steps 1 --warmup 1 2>&1 >/dev/null | grep "\[plan\]" | grep "branches.2" | sort -u | cut -c1-200 | head -8 > gpurun_out/${T}_plans_b2.log
