#!/bin/bash
mkdir -p gpurun_out
T=${1:-r02t}
timeout 900 python -m pytest tests/test_gpu_hrnet_parity.py tests/test_gpu_parity.py -m gpu -x -q --timeout 600 -k "hrnet or pipeline or conv" 2>&1 | tail -3 | tee gpurun_out/${T}_pytest.log
run() { lbl=$1; shift
  env "$@" timeout 300 python bench.py --steps 10 --warmup 3 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('$lbl hrnet_ms', round(d['roofline']['hrnet_ms'],4), 'parity', d.get('parity_check','')[:10])" | tee -a gpurun_out/${T}_variants.log
}
run teams3_n64 HBP_X=0
run teams2 HBP_HALO_TEAMS=2
run teams3_n32 HBP_HALO_TEAMS3_N=32
run teams3_n128 HBP_HALO_TEAMS3_N=128
run teams2_again HBP_HALO_TEAMS=2
for t in 2 3; do
echo "== teams $t" | tee -a gpurun_out/${T}_mb.log
HBP_HALO_TEAMS=$t HBP_HALO_TEAMS3_N=32 HBP_MB_SHAPES=0,1,2,4,5,7 timeout 120 python tools/conv_microbench.py 2>/dev/null | grep "eng=" | tee -a gpurun_out/${T}_mb.log
done
