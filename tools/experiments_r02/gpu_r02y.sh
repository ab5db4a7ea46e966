#!/bin/bash
mkdir -p gpurun_out
T=${1:-r02y}
run() { lbl=$1; shift
  env "$@" timeout 300 python bench.py --steps 10 --warmup 3 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('$lbl hrnet_ms', round(d['roofline']['hrnet_ms'],4), 'parity', d.get('parity_check','')[:10])" | tee -a gpurun_out/${T}_variants.log
}
run base HBP_X=0
run s3_42_24_34 HBP_BRANCH_SHARE3=0.42,0.24,0.34
run s3_40_26_34 HBP_BRANCH_SHARE3=0.40,0.26,0.34
run s3_44_24_32 HBP_BRANCH_SHARE3=0.44,0.24,0.32
run s3_40_24_36 HBP_BRANCH_SHARE3=0.40,0.24,0.36
run s3_46_22_32 HBP_BRANCH_SHARE3=0.46,0.22,0.32
run s3_42_22_36 HBP_BRANCH_SHARE3=0.42,0.22,0.36
run s4_34_22_26_18 HBP_BRANCH_SHARE4=0.34,0.22,0.26,0.18
run s4_36_22_24_18 HBP_BRANCH_SHARE4=0.36,0.22,0.24,0.18
run s4_34_20_26_20 HBP_BRANCH_SHARE4=0.34,0.20,0.26,0.20
run s4_36_18_26_20 HBP_BRANCH_SHARE4=0.36,0.18,0.26,0.20
run s4_38_20_24_18 HBP_BRANCH_SHARE4=0.38,0.20,0.24,0.18
run s2_60_40 HBP_BRANCH_SHARE2=0.60,0.40
run s2_68_32 HBP_BRANCH_SHARE2=0.68,0.32
