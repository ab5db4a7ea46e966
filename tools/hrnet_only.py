"""A few HRNet-W32 forwards at 64 crops, device resident (the process ncu captures the conv kernels from).
usage: python tools/hrnet_only.py [forwards]"""
import ctypes as C
import sys

import numpy as np

sys.path.insert(0, ".")
from human_body_proportion_estimation_b200 import _capi, engine as E  # noqa: E402

eng = E.Engine(0)
eng.load_hrnet(None, 32, 256, 192, seed=0)
P = 64
x = np.random.default_rng(0).uniform(0, 1, (P, 3, 256, 192)).astype(np.float16)
d_x = eng.to_device(x)
d_hm = eng.dev_alloc(P * 17 * 64 * 48 * 2)
for _ in range(int(sys.argv[1]) if len(sys.argv) > 1 else 3):
    E.check(eng._lib.hbp_hrnet_forward(eng._ctx, C.c_void_p(d_x), P, C.c_void_p(d_hm), _capi.F16, _capi.DEVICE))
eng.sync()
print("ok")
