#!/usr/bin/env python
"""Batched K4 launch alone (4096 crops from 64 4K frames -> (4096,3,256,192) f16): the launch bench.py's
stage roofline times, for `ncu -k regex:crop_warp`.  Prints the CUDA-event time of the best of 5."""
import ctypes as C
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from human_body_proportion_estimation_b200 import geometry, synth
from human_body_proportion_estimation_b200._capi import DEVICE, F16, check
from human_body_proportion_estimation_b200.engine import Engine

eng = Engine(0)
lib, ctx = eng._lib, eng._ctx
nf, fh, fw, P, IN_H, IN_W = 64, 2160, 3840, 4096, 256, 192
frame = synth.frame_u8(fh, fw, seed=synth.SEED_BASE + 5, smooth=False)
d_frames = eng.dev_alloc(nf * frame.nbytes)
for i in range(nf):
    eng.h2d(d_frames + i * frame.nbytes, frame)
boxes = synth.person_boxes_yxyx_px(P, fh, fw, seed=synth.SEED_BASE + 6, hmin=300, hmax=1400)
mats = geometry.crop_and_resize_matrices(boxes / np.array([fh, fw, fh, fw], np.float32), fh, fw, IN_H, IN_W)
fidx = (np.arange(P) // 64).astype(np.int32)
d_m = eng.to_device(mats.reshape(P, 6)); d_fi = eng.to_device(fidx)
d_cr = eng.dev_alloc(P * 3 * IN_H * IN_W * 2)
best = 1e9
for _ in range(int(os.environ.get("REPS", "5"))):
    eng.flush_l2()
    eng.timer_start(3)
    check(lib.hbp_crop_warp(ctx, C.c_void_p(d_frames), nf, fh, fw, C.c_void_p(d_m), C.c_void_p(d_fi), P, IN_H, IN_W, 1,
                            C.c_void_p(d_cr), F16, DEVICE))
    eng.timer_stop(3)
    best = min(best, eng.timer_ms(3))
bw = boxes[:, 3] - boxes[:, 1]; bh = boxes[:, 2] - boxes[:, 0]
print("crop 4096: %.3f ms; mean box %.0f x %.0f px; box bytes %.2f GB, out %.2f GB" % (
    best, bw.mean(), bh.mean(), float((bw * bh).sum()) * 3 / 1e9, P * 3 * IN_H * IN_W * 2 / 1e9))
