#!/usr/bin/env python
"""Per-launch device time and TFLOP/s of the HRNet-W32 conv shape classes at batch 64
(hbp_conv2d_nhwc_timed: one plan, N launches between two CUDA events)."""
import json
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
# (H, W, Cin, Cout, k, stride, up, res, share of W32 FLOPs in %)
SHAPES = [
    (64, 48, 32, 32, 3, 1, 1, True, 23.7), (32, 24, 64, 64, 3, 1, 1, True, 23.7),
    (16, 12, 128, 128, 3, 1, 1, True, 20.7), (8, 6, 256, 256, 3, 1, 1, True, 8.9),
    (64, 48, 64, 64, 3, 1, 1, False, 5.9), (64, 48, 64, 256, 1, 1, 1, True, 3.3),
    (64, 48, 256, 32, 3, 1, 1, False, 3.0), (64, 48, 256, 64, 1, 1, 1, False, 2.0),
    (128, 96, 64, 64, 3, 2, 1, False, 1.5), (64, 48, 256, 64, 3, 2, 1, False, 1.5),
    (64, 48, 32, 64, 3, 2, 1, True, 1.0), (32, 24, 64, 32, 1, 1, 2, True, 0.1),
    (8, 6, 256, 32, 1, 1, 8, True, 0.1),
]


def main():
    from human_body_proportion_estimation_b200.engine import Engine
    P = int(os.environ.get("HBP_MB_BATCH", "64"))
    engine = int(os.environ.get("HBP_MB_ENGINE", "1"))
    eng = Engine(0)
    rng = np.random.default_rng(0)
    rows = []
    sel = os.environ.get("HBP_MB_SHAPES")
    shapes = [SHAPES[int(i)] for i in sel.split(",")] if sel else SHAPES
    for H, W, Cin, Cout, k, s, up, use_res, share in shapes:
        x = rng.standard_normal((P, H, W, Cin)).astype(np.float16)
        w = (rng.standard_normal((Cout, Cin, k, k)) / np.sqrt(Cin * k * k)).astype(np.float16)
        b = np.zeros(Cout, np.float32)
        if os.environ.get("HBP_MB_NORES"):
            use_res = False
        res = rng.standard_normal((P, H // s * up, W // s * up, Cout)).astype(np.float16) if use_res else None
        iters = int(os.environ.get("HBP_MB_ITERS", "50"))
        r = eng.conv2d_nhwc(x, w, b, res, s, up, True, engine, time_iters=iters)
        if iters <= 0:            # HBP_CONV_TRACE runs: plan + per-CTA phase trace on stderr only
            continue
        _, used, ms = r
        flop = 2.0 * P * (H // s) * (W // s) * Cin * Cout * k * k
        rows.append(dict(shape=[H, W, Cin, Cout, k, s, up], engine=used, us=ms * 1e3, tflops=flop / ms / 1e9,
                         share_pct=share))
        print("%-34s eng=%d %8.2f us %8.1f TFLOP/s  (%.1f %% of W32 FLOPs)" %
              ((H, W, Cin, Cout, k, s, up), used, ms * 1e3, flop / ms / 1e9, share), flush=True)
    if not rows:
        return
    tot = sum(r["share_pct"] for r in rows)
    t = sum(r["share_pct"] / r["tflops"] for r in rows)
    print("FLOP-weighted harmonic mean over %.1f %% of the network: %.1f TFLOP/s" % (tot, tot / t))
    json.dump(rows, open(os.path.join("gpurun_out", "conv_microbench.json"), "w"), indent=1)


if __name__ == "__main__":
    main()
