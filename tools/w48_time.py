#!/usr/bin/env python
"""HRNet-W48 384x288 forward timing (BASELINE configs[3] shape), device-resident, CUDA events."""
import ctypes as C, os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from human_body_proportion_estimation_b200 import hrnet_arch
from human_body_proportion_estimation_b200._capi import DEVICE, F16, check
from human_body_proportion_estimation_b200.engine import Engine
eng = Engine(0)
eng.load_hrnet(None, 48, 384, 288, seed=0)
lib, ctx = eng._lib, eng._ctx
flops, _ = hrnet_arch.flops_per_crop(48, 384, 288)
for P in (16, 64):
    crops = np.random.default_rng(0).random((P, 3, 384, 288), dtype=np.float32).astype(np.float16)
    d_c = eng.to_device(crops); d_h = eng.dev_alloc(P * 17 * 96 * 72 * 2)
    for _ in range(3):
        check(lib.hbp_hrnet_forward(ctx, C.c_void_p(d_c), P, C.c_void_p(d_h), F16, DEVICE))
    eng.sync()
    ms = []
    for _ in range(5):
        eng.flush_l2(); eng.timer_start(3)
        check(lib.hbp_hrnet_forward(ctx, C.c_void_p(d_c), P, C.c_void_p(d_h), F16, DEVICE))
        eng.timer_stop(3); ms.append(eng.timer_ms(3))
    t = min(ms)
    print("W48 384x288 P=%d: %.3f ms  %.1f crops/s  %.1f TFLOP/s (algorithmic %.2f GFLOP/crop)" % (P, t, P / t * 1e3, flops * P / t / 1e9, flops / 1e9))
    eng.dev_free(d_c); eng.dev_free(d_h)
