#!/bin/bash
mkdir -p gpurun_out
T=${1:-r02u}
run() { lbl=$1; shift
  env "$@" timeout 300 python bench.py --steps 10 --warmup 3 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('$lbl hrnet_ms', round(d['roofline']['hrnet_ms'],4), 'parity', d.get('parity_check','')[:10])" | tee -a gpurun_out/${T}_variants.log
}
run base HBP_HALO_TEAMS=2
run m1 HBP_HALO_TEAMS=2 HBP_HALO_M=1
run l2bpc32 HBP_HALO_TEAMS=2 HBP_HALO_L2BPC=32
run l2bpc128 HBP_HALO_TEAMS=2 HBP_HALO_L2BPC=128
run astages2 HBP_HALO_TEAMS=2 HBP_HALO_ASTAGES=2
run issuers1 HBP_HALO_TEAMS=2 HBP_HALO_ISSUERS=1
HBP_HALO_TEAMS=2 HBP_CONV_TRACE=1 timeout 120 python bench.py --steps 1 --warmup 1 2>&1 >/dev/null | grep "\[plan\]\|\[pgroup\]" | sort -u > gpurun_out/${T}_plans.log
