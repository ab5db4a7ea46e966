"""End-to-end chained det->pose step (configs[1]: one 1080p frame, 64 persons) from pinned host buffers:
one Engine with two tickets in flight  vs  two Engines (two hbp contexts: own streams, buffers, HRNet weights and graphs)
on the SAME GPU, frames alternating between them.  usage: python tools/dual_engine_e2e.py [steps]"""
import sys
import time

import numpy as np

sys.path.insert(0, ".")
from human_body_proportion_estimation_b200 import engine as E, synth  # noqa: E402

steps = int(sys.argv[1]) if len(sys.argv) > 1 else 100
FH, FW = 1080, 1920
frame = np.stack([synth.frame_u8(FH, FW, seed=synth.SEED_BASE + 2)])
pred = synth.yolo_head_grid()[0]
HEIGHTS = [175.0, 168.0, 181.0]


def make(n_eng):
    engs = [E.Engine(0) for _ in range(n_eng)]
    w = None
    for e in engs:
        w = e.load_hrnet(w, 32, 256, 192, seed=0)
    bufs = []
    for e in engs:
        per = []
        for _ in range(2):
            f = e.pinned_empty(frame.shape, np.uint8); f[...] = frame
            p = e.pinned_empty(pred.shape, np.float32); p[...] = pred
            per.append((f, p))
        bufs.append(per)
    return engs, bufs


def run(n_eng, depth):
    engs, bufs = make(n_eng)

    def submit(i):
        e = engs[i % n_eng]
        f, p = bufs[i % n_eng][(i // n_eng) % 2]
        return e, e.det_pose_submit_yolo(f, p, person_height=HEIGHTS, persons_cap=64, resample="bilinear")

    for i in range(6):                          # warm-up: plans, graphs
        e, tk = submit(i)
        e.det_pose_collect(tk)
    inflight, crops = [], 0
    t0 = time.perf_counter()
    for i in range(steps):
        inflight.append(submit(i))
        if len(inflight) >= depth * n_eng:
            e, tk = inflight.pop(0)
            crops += e.det_pose_collect(tk)["n"]
    while inflight:
        e, tk = inflight.pop(0)
        crops += e.det_pose_collect(tk)["n"]
    dt = time.perf_counter() - t0
    print("engines=%d tickets/engine=%d: %.0f crops/s end to end (%.3f ms per 64-person frame)" % (n_eng, depth, crops / dt, dt / steps * 1e3), flush=True)
    for e in engs:
        e.close()


if __name__ == "__main__":
    run(1, 2)
    run(2, 1)
    run(2, 2)
    run(3, 1)
