#!/bin/bash
mkdir -p gpurun_out
T=r02ak
run() { lbl=$1; shift
  env "$@" timeout 300 python bench.py --steps 10 --warmup 3 2>gpurun_out/${T}_$lbl.err | python -c "import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('$lbl hrnet_ms', round(d['roofline']['hrnet_ms'],4), 'parity', d.get('parity_check','')[:10])" | tee -a gpurun_out/${T}_variants.log
}
run split1_a HBP_X=0
run split0_a HBP_HALO_SPLIT_PRODUCER=0
run split1_b HBP_X=0
run split0_b HBP_HALO_SPLIT_PRODUCER=0
for sp in 1 0; do
echo "== split $sp" | tee -a gpurun_out/${T}_variants.log
HBP_HALO_SPLIT_PRODUCER=$sp HBP_MB_SHAPES=6,2 timeout 100 python tools/conv_microbench.py 2>&1 | grep "eng=" | tee -a gpurun_out/${T}_variants.log
done
timeout 900 python -m pytest tests -m gpu -x -q --timeout 600 2>&1 | tail -3 | tee gpurun_out/${T}_pytest.log
