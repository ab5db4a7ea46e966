#!/bin/bash
mkdir -p gpurun_out
T=r02aj
run() { lbl=$1; shift
  env "$@" timeout 300 python bench.py --steps 10 --warmup 3 2>gpurun_out/${T}_$lbl.err | python -c "import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('$lbl hrnet_ms', round(d['roofline']['hrnet_ms'],4), 'parity', d.get('parity_check','')[:10])" | tee -a gpurun_out/${T}_variants.log
}
run base HBP_X=0
run pgm2 HBP_PG_M2=1
run pgm2_s4a HBP_PG_M2=1 HBP_BRANCH_SHARE4=0.38,0.22,0.24,0.16
run pgm2_s4b HBP_PG_M2=1 HBP_BRANCH_SHARE4=0.40,0.20,0.24,0.16
run s4c HBP_BRANCH_SHARE4=0.36,0.20,0.24,0.20
HBP_PG_M2=1 HBP_MB_SHAPES=3 HBP_HALO_SMS=27 timeout 100 python tools/conv_microbench.py 2>&1 | grep "eng=\|pgroup\]" | cut -c1-160 | tee -a gpurun_out/${T}_variants.log
HBP_MB_SHAPES=3 HBP_HALO_SMS=27 timeout 100 python tools/conv_microbench.py 2>&1 | grep "eng=\|pgroup\]" | cut -c1-160 | tee -a gpurun_out/${T}_variants.log
