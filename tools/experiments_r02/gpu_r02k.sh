#!/bin/bash
mkdir -p gpurun_out
HBP_CHAIN_TRACE=stage3.1.branches.0 timeout 300 python - > gpurun_out/r02k_chaintrace.log 2>&1 <<'PY'
import numpy as np
from human_body_proportion_estimation_b200.engine import Engine
e = Engine(0)
e.load_hrnet(None, 32, 256, 192, seed=0)
x = np.random.default_rng(0).uniform(0, 1, (64, 3, 256, 192)).astype(np.float16)
for i in range(3):
    a = e.hrnet_forward(x)
print("done")
PY
grep -c chaintrace gpurun_out/r02k_chaintrace.log
