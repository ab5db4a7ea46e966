#!/bin/bash
mkdir -p gpurun_out
T=${1:-r02d}
timeout 300 python tools/bringup_conv.py > gpurun_out/${T}_bringup.log 2>&1; echo "bringup rc=$?"
grep -c ok gpurun_out/${T}_bringup.log; grep -v " ok" gpurun_out/${T}_bringup.log | head
for P in 64 256; do
  HBP_MB_BATCH=$P HBP_MB_SHAPES=0,1,2,3 timeout 200 python tools/conv_microbench.py 2>&1 | grep eng= | sed "s/^/P=$P /" | tee -a gpurun_out/${T}_mb.log
done
HBP_HALO_RES_MMA=0 HBP_MB_BATCH=256 HBP_MB_SHAPES=0,1 timeout 200 python tools/conv_microbench.py 2>&1 | grep eng= | sed "s/^/nores_mma P=256 /" | tee -a gpurun_out/${T}_mb.log
HBP_HALO_BIAS_MMA=0 HBP_MB_BATCH=256 HBP_MB_SHAPES=0,1 timeout 200 python tools/conv_microbench.py 2>&1 | grep eng= | sed "s/^/nobias_mma P=256 /" | tee -a gpurun_out/${T}_mb.log
timeout 1200 python -m pytest tests -m gpu -x -q --timeout 900 2>&1 | tail -15 | tee gpurun_out/${T}_pytest.log
timeout 600 python bench.py --steps 10 --warmup 3 > gpurun_out/${T}_bench.json 2> gpurun_out/${T}_bench.err; echo "bench rc=$?"
python - <<PY
import json
d = json.load(open("gpurun_out/${T}_bench.json"))
print("value", d["value"], "e2e", d["e2e"]["value"], "hrnet_ms", d["roofline"]["hrnet_ms"], "frac", d["roofline"]["frac"])
PY
