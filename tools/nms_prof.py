"""One-frame NMS (25200 x 85 head, 64 persons kept) and the 128-head candidate filter, device resident: per-call time with
CUDA events; run under `ncu --metrics gpu__time_duration.sum` for the per-kernel split."""
import ctypes as C
import sys

import numpy as np

sys.path.insert(0, ".")
from human_body_proportion_estimation_b200 import _capi, engine as E, synth  # noqa: E402

DEVICE = _capi.DEVICE


def main():
    eng = E.Engine(0)
    lib, ctx = eng._lib, eng._ctx
    pred, _ = synth.yolo_head_grid()
    d_pred = eng.to_device(pred)
    d_cls = eng.to_device(np.zeros(1, np.int32))
    d_det = eng.dev_alloc(300 * 6 * 4); d_cnt = eng.dev_alloc(4)

    def nms():
        E.check(lib.hbp_yolo_nms(ctx, C.c_void_p(d_pred), 1, 25200, 80, 0.4, 0.5, C.c_void_p(d_cls), 1, 300, C.c_void_p(d_det),
                                 C.c_void_p(d_cnt), DEVICE))
    reps = int(sys.argv[1]) if len(sys.argv) > 1 else 20
    nms(); eng.sync()
    eng.timer_start(2)
    for _ in range(reps):
        nms()
    eng.timer_stop(2)
    cnt = np.zeros(1, np.int32); eng.d2h(cnt, d_cnt); eng.sync()
    print("hbp_yolo_nms 1 x 25200 x 85: %.1f us per call, kept %d (host enqueues while the GPU runs: max(host, device))" % (eng.timer_ms(2) / reps * 1e3, cnt[0]), flush=True)
    # device time alone: the calls are enqueued behind ~2.4 ms of other work on the stream (an HRNet forward), so the GPU
    # never waits for the host inside the timed region
    eng.load_hrnet(None, 32, 256, 192, seed=0)
    x = np.zeros((64, 3, 256, 192), np.float16)
    d_x = eng.to_device(x); d_hm = eng.dev_alloc(64 * 17 * 64 * 48 * 2)
    for _ in range(3):
        E.check(lib.hbp_hrnet_forward(ctx, C.c_void_p(d_x), 64, C.c_void_p(d_hm), _capi.F16, DEVICE))
    eng.sync()
    E.check(lib.hbp_hrnet_forward(ctx, C.c_void_p(d_x), 64, C.c_void_p(d_hm), _capi.F16, DEVICE))
    eng.timer_start(2)
    for _ in range(reps):
        nms()
    eng.timer_stop(2)
    eng.sync()
    print("hbp_yolo_nms 1 x 25200 x 85: %.1f us per call on the device (queued behind an HRNet forward)" % (eng.timer_ms(2) / reps * 1e3), flush=True)
    B = 128
    head = synth.yolo_decoded_head()[0]
    d_big = eng.dev_alloc(B * head.nbytes)
    for i in range(B):
        eng.h2d(d_big + i * head.nbytes, head)
    d_c = eng.dev_alloc(B * 4)

    def filt():
        E.check(lib.hbp_yolo_filter(ctx, C.c_void_p(d_big), B, 25200, 80, 0.4, C.c_void_p(d_cls), 1, 4096, C.c_void_p(d_c), DEVICE))
    filt(); eng.sync()
    eng.timer_start(2)
    for _ in range(max(2, reps // 4)):
        filt()
    eng.timer_stop(2)
    print("hbp_yolo_filter 128 x 25200 x 85: %.1f us per call" % (eng.timer_ms(2) / max(2, reps // 4) * 1e3), flush=True)


if __name__ == "__main__":
    main()
