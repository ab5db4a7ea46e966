#!/usr/bin/env python
"""PIL-bicubic letterbox batch alone (64 4K frames -> 3x640x640 f16), for `ncu -k regex:pil_`."""
import ctypes as C, os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from human_body_proportion_estimation_b200 import synth
from human_body_proportion_estimation_b200._capi import DEVICE, F16, NCHW, PRE_LETTERBOX_PIL, check
from human_body_proportion_estimation_b200.engine import Engine
eng = Engine(0)
lib, ctx = eng._lib, eng._ctx
nf, fh, fw = 64, 2160, 3840
frame = synth.frame_u8(fh, fw, seed=synth.SEED_BASE + 5, smooth=False)
d_frames = eng.dev_alloc(nf * frame.nbytes)
for i in range(nf):
    eng.h2d(d_frames + i * frame.nbytes, frame)
d_lb = eng.dev_alloc(nf * 3 * 640 * 640 * 2)
best = 1e9
for _ in range(int(os.environ.get("REPS", "5"))):
    eng.flush_l2(); eng.timer_start(3)
    check(lib.hbp_preprocess(ctx, C.c_void_p(d_frames), nf, fh, fw, PRE_LETTERBOX_PIL, 640, 640, 1, 128, C.c_void_p(d_lb), F16, NCHW, DEVICE))
    eng.timer_stop(3); best = min(best, eng.timer_ms(3))
print("PIL letterbox 64x4K: %.3f ms" % best)
