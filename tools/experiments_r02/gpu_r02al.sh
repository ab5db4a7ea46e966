#!/bin/bash
mkdir -p gpurun_out
T=r02al
run() { lbl=$1; shift
  env "$@" timeout 300 python bench.py --steps 10 --warmup 3 2>gpurun_out/${T}_$lbl.err | python -c "import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('$lbl hrnet_ms', round(d['roofline']['hrnet_ms'],4), 'parity', d.get('parity_check','')[:10])" | tee -a gpurun_out/${T}_variants.log
}
run base HBP_X=0
run c64m1_res HBP_HALO_C64_M1=1
run c64m1_all HBP_HALO_C64_M1=2
run base2 HBP_X=0
run c64m1_res_s3 HBP_HALO_C64_M1=1 HBP_BRANCH_SHARE3=0.44,0.22,0.34
run c64m1_res_s3b HBP_HALO_C64_M1=1 HBP_BRANCH_SHARE3=0.46,0.22,0.32
HBP_HALO_C64_M1=1 HBP_CONV_TRACE=1 timeout 120 python bench.py --steps 1 --warmup 1 2>&1 >/dev/null | grep "\[plan\] stage3.1.branches.1.0" | cut -c1-200 | sort -u | tee -a gpurun_out/${T}_variants.log
