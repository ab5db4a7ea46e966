#!/bin/bash
mkdir -p gpurun_out
T=${1:-r02v}
run() { lbl=$1; shift
  env "$@" timeout 300 python bench.py --steps 10 --warmup 3 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('$lbl hrnet_ms', round(d['roofline']['hrnet_ms'],4), 'parity', d.get('parity_check','')[:10])" | tee -a gpurun_out/${T}_variants.log
}
run bpsm8 HBP_X=0
run bpsm1 HBP_UPADD_BPSM=1
run bpsm2 HBP_UPADD_BPSM=2
run bpsm3 HBP_UPADD_BPSM=3
run bpsm4 HBP_UPADD_BPSM=4
HBP_UPADD_BPSM=1 HBP_TIMELINE=1 timeout 300 python bench.py --steps 3 --warmup 3 > /dev/null 2> gpurun_out/${T}_timeline_bpsm1.log
HBP_UPADD_BPSM=2 HBP_TIMELINE=1 timeout 300 python bench.py --steps 3 --warmup 3 > /dev/null 2> gpurun_out/${T}_timeline_bpsm2.log
