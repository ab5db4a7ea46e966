#!/bin/bash
mkdir -p gpurun_out
timeout 300 python - > gpurun_out/r02i_dbg.log 2>&1 <<'PY'
import numpy as np, traceback
from human_body_proportion_estimation_b200.engine import Engine
e = Engine(0)
e.load_hrnet(None, 32, 256, 192, seed=0)
x = np.random.default_rng(0).uniform(0, 1, (64, 3, 256, 192)).astype(np.float16)
try:
    a = e.hrnet_forward(x)
    b = e.hrnet_forward(x)
    c = e.hrnet_forward(x)
    print("ok", np.array_equal(a, b), np.array_equal(b, c), float(np.abs(a).max()))
except Exception as ex:
    traceback.print_exc()
PY
tail -30 gpurun_out/r02i_dbg.log
