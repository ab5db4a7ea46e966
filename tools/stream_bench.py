#!/usr/bin/env python
"""BASELINE configs[4]: synthetic 4K frame stream, ~100 persons per frame, frames sharded over the GPUs of the box
through the product API (MultiGpuEngine.stream: frame f -> GPU f mod G, two frames in flight per GPU, no
collective).  Prints one JSON line: crops/s, frames/s, p50/p95 frame latency.

    python tools/stream_bench.py --gpus 2 --frames 400 --warmup 40 [--persons 100]
"""
import argparse
import json
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from human_body_proportion_estimation_b200 import geometry, synth                      # noqa: E402
from human_body_proportion_estimation_b200.engine import MultiGpuEngine               # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--frames", type=int, default=400)
    ap.add_argument("--warmup", type=int, default=40)
    ap.add_argument("--persons", type=int, default=100)
    args = ap.parse_args()
    H, W = 2160, 3840
    mg = MultiGpuEngine(list(range(args.gpus)), width=32, in_h=256, in_w=192, seed=0)
    # 8 distinct pre-staged pinned frames per GPU, cycled; 100 boxes per frame, h in [150, 600]
    base = [synth.frame_u8(H, W, seed=synth.SEED_BASE + 50 + i, smooth=False) for i in range(8)]
    pinned = []
    for eng in mg.engines:
        bufs = []
        for b in base:
            p = eng.pinned_empty(b.shape, np.uint8)
            p[...] = b
            bufs.append(p)
        pinned.append(bufs)
    sets = []
    for i in range(8):
        boxes = synth.person_boxes_yxyx_px(args.persons, H, W, seed=synth.SEED_BASE + 60 + i, hmin=150, hmax=600)
        mats = geometry.crop_and_resize_matrices(boxes / np.array([H, W, H, W], np.float32), H, W, 256, 192)
        sets.append((mats.reshape(-1, 6), boxes))
    G = args.gpus

    def source(f):
        mats, boxes = sets[f % 8]
        return pinned[f % G][(f // G) % 8], mats, boxes, 175.0

    mg.stream(source, args.warmup)
    t0 = time.perf_counter()
    res, lat = mg.stream(source, args.frames)
    dt = time.perf_counter() - t0
    assert all(r is not None and r["kpts_img"].shape == (args.persons, 17, 2) for r in res)
    lat = np.sort(np.asarray(lat))
    print(json.dumps({
        "workload": "configs[4]: 4K frames (2160x3840x3 u8, pinned), %d persons/frame, HRNet-W32 256x192, frame f -> GPU f mod G" % args.persons,
        "n_gpus": G, "frames": args.frames, "warmup_frames": args.warmup,
        "crops_per_s": args.frames * args.persons / dt, "frames_per_s": args.frames / dt,
        "p50_frame_latency_ms": float(lat[len(lat) // 2]), "p95_frame_latency_ms": float(lat[int(len(lat) * 0.95)]),
        "h2d_bytes_per_frame": int(H * W * 3), "api": "MultiGpuEngine.stream (hbp_pose_pipeline_submit/_collect, 2 frames in flight per GPU)"}))


if __name__ == "__main__":
    main()
