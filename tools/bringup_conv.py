#!/usr/bin/env python
"""Bring-up / regression of the fused conv operator (hbp_conv2d_nhwc) on a GPU box:
every HRNet shape class through engine 1 (tcgen05/TMA) and engine 0 (SIMT) against a
torch fp32 convolution of the same fp16-rounded operands.

    python tools/bringup_conv.py [--engine 1] [--cases 0,3,11]

Each case prints max|err| / max|ref| ; exits non-zero if any case exceeds 4e-3.
"""
import argparse
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

# (P, H, W, Cin, Cout, k, stride, up, residual, relu)
CASES = [
    (2, 64, 48, 64, 64, 1, 1, 1, False, True),      # layer1 1x1
    (2, 64, 48, 64, 64, 3, 1, 1, False, True),      # layer1 3x3, SW128
    (3, 64, 48, 32, 32, 3, 1, 1, True, True),       # branch 0 BasicBlock conv2, SW64
    (3, 32, 24, 64, 64, 3, 1, 1, True, True),       # branch 1
    (3, 16, 12, 128, 128, 3, 1, 1, True, True),     # branch 2 (2-image tiles)
    (9, 8, 6, 256, 256, 3, 1, 1, True, True),       # branch 3 (8-image tiles, ragged batch)
    (2, 64, 48, 256, 64, 1, 1, 1, False, True),     # bottleneck conv1
    (2, 64, 48, 64, 256, 1, 1, 1, True, True),      # bottleneck conv3 + residual
    (2, 64, 48, 256, 32, 3, 1, 1, False, True),     # transition1.0
    (2, 32, 24, 64, 32, 1, 1, 2, True, False),      # fuse up x2
    (2, 8, 6, 256, 32, 1, 1, 8, True, True),        # fuse up x8
    (2, 128, 96, 64, 64, 3, 2, 1, False, True),     # stem conv2 (stride 2)
    (2, 64, 48, 32, 64, 3, 2, 1, True, True),       # fuse down
    (2, 64, 48, 256, 64, 3, 2, 1, False, True),     # transition1.1
    (2, 16, 12, 128, 256, 3, 2, 1, True, False),    # fuse down to branch 3
    (64, 64, 48, 32, 32, 3, 1, 1, True, True),      # full batch
    (3, 24, 18, 192, 192, 3, 1, 1, True, True),     # W48 branch 2 (halo rows 24 -> 75 % tiles)
    (5, 12, 9, 384, 384, 3, 1, 1, True, False),     # W48 branch 3 (stacked images, ragged width)
    (64, 32, 24, 64, 64, 3, 1, 1, True, True),      # full batch branch 1
    (64, 16, 12, 128, 128, 3, 1, 1, True, True),    # full batch branch 2
    (64, 8, 6, 256, 256, 3, 1, 1, True, True),      # full batch branch 3
    # HRNet-W48 384x288 shape classes (48/96 channels are stored padded to 64/128); widths 72/36/18/9
    (2, 96, 72, 64, 64, 3, 1, 1, True, True),       # 21 W48 branch 0
    (2, 48, 36, 128, 128, 3, 1, 1, True, True),     # 22 W48 branch 1 (ragged width 36)
    (2, 96, 72, 256, 64, 3, 1, 1, False, True),     # 23 W48 transition1.0
    (2, 96, 72, 256, 128, 3, 2, 1, False, True),    # 24 W48 transition1.1
    (2, 96, 72, 64, 128, 3, 2, 1, True, True),      # 25 W48 fuse 0->1
    (2, 48, 36, 128, 192, 3, 2, 1, True, True),     # 26 W48 fuse 1->2 (output width 18)
    (2, 24, 18, 192, 384, 3, 2, 1, True, False),    # 27 W48 fuse 2->3 (output width 9)
    (2, 96, 72, 64, 64, 3, 2, 1, False, True),      # 28 W48 fuse 0->2/3 first link
    (2, 48, 36, 128, 64, 1, 1, 1, False, False),    # 29 W48 fuse 1x1 1->0
    (2, 24, 18, 192, 128, 1, 1, 1, False, False),   # 30 W48 fuse 1x1 2->1
    (2, 12, 9, 384, 192, 1, 1, 1, False, False),    # 31 W48 fuse 1x1 3->2
    (2, 12, 9, 384, 64, 1, 1, 1, False, False),     # 32 W48 fuse 1x1 3->0
]


def reference(x, w, b, res, stride, up, relu):
    """fp32 convolution by im2col + one matmul (numpy only: no torch import on the box)"""
    P, H, W, Cin = x.shape
    Cout, _, k, _ = w.shape
    pad = k // 2
    xp = np.zeros((P, H + 2 * pad, W + 2 * pad, Cin), np.float32)
    xp[:, pad:pad + H, pad:pad + W] = x
    Ho, Wo = H // stride, W // stride
    cols = np.empty((P, Ho, Wo, k * k * Cin), np.float32)
    for dy in range(k):
        for dx in range(k):
            t = dy * k + dx
            cols[..., t * Cin:(t + 1) * Cin] = xp[:, dy:dy + Ho * stride:stride, dx:dx + Wo * stride:stride]
    wm = np.transpose(w.astype(np.float32), (2, 3, 1, 0)).reshape(k * k * Cin, Cout)
    y = cols.reshape(-1, k * k * Cin) @ wm + b
    y = y.reshape(P, Ho, Wo, Cout)
    if up > 1:
        y = y.repeat(up, axis=1).repeat(up, axis=2)
    if res is not None:
        y = y + res.astype(np.float32)
    return np.maximum(y, 0) if relu else y


def run_case(eng, idx, engine):
    P, H, W, Cin, Cout, k, s, up, use_res, relu = CASES[idx]
    rng = np.random.default_rng(100 + idx)
    x = rng.standard_normal((P, H, W, Cin)).astype(np.float16)
    w = (rng.standard_normal((Cout, Cin, k, k)) / np.sqrt(Cin * k * k)).astype(np.float16)
    b = rng.standard_normal(Cout).astype(np.float32) * 0.1
    res = rng.standard_normal((P, H // s * up, W // s * up, Cout)).astype(np.float16) if use_res else None
    out, used = eng.conv2d_nhwc(x, w, b, res, s, up, relu, engine)
    ref = reference(x, w, b, res, s, up, relu)
    err = np.abs(out.astype(np.float32) - ref).max() / np.abs(ref).max()
    return err, used


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--engine", type=int, default=1)
    ap.add_argument("--cases", type=str, default="")
    a = ap.parse_args()
    from human_body_proportion_estimation_b200.engine import Engine
    eng = Engine(0)
    bad = 0
    for i in range(len(CASES)):
        if a.cases and str(i) not in a.cases.split(","):
            continue
        try:
            err, used = run_case(eng, i, a.engine)
            ok = err < 4e-3
            print("case %2d %-46s engine=%d err=%.3e %s" % (i, CASES[i], used, err, "ok" if ok else "FAIL"), flush=True)
            bad += 0 if ok else 1
        except Exception as e:
            print("case %2d %-46s EXC %s" % (i, CASES[i], e), flush=True)
            bad += 1
            break           # a CUDA fault poisons the context
    sys.exit(1 if bad else 0)


if __name__ == "__main__":
    main()
