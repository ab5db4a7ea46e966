#!/bin/bash
mkdir -p gpurun_out
T=${1:-r02s}
timeout 900 python -m pytest tests/test_gpu_hrnet_parity.py tests/test_gpu_parity.py -m gpu -x -q --timeout 600 -k "hrnet or pipeline or conv" 2>&1 | tail -3 | tee gpurun_out/${T}_pytest.log
run() { lbl=$1; shift
  env "$@" timeout 300 python bench.py --steps 10 --warmup 3 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('$lbl hrnet_ms', round(d['roofline']['hrnet_ms'],4), 'parity', d.get('parity_check','')[:10])" | tee -a gpurun_out/${T}_variants.log
}
run base HBP_X=0
run nostores HBP_HALO_DBG=1
run noepilogue HBP_HALO_DBG=8
HBP_MB_SHAPES=0,1,2,4,5,7 timeout 120 python tools/conv_microbench.py 2>/dev/null | grep "eng=" | tee -a gpurun_out/${T}_mb.log
HBP_MB_BATCH=256 HBP_MB_SHAPES=0,1,2 timeout 120 python tools/conv_microbench.py 2>/dev/null | grep "eng=" | tee -a gpurun_out/${T}_mb.log
