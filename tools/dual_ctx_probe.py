"""Does running the HRNet forward as K independent sub-batches on K contexts (own streams, one CUDA context) hide the
per-level fixed cost?  Times total crops/s for (K contexts x P crops each), device-resident inputs, CUDA-graph replays.
usage: python tools/dual_ctx_probe.py  (prints one line per configuration)"""
import ctypes as C
import sys
import time

import numpy as np

sys.path.insert(0, ".")
from human_body_proportion_estimation_b200 import _capi, engine as E  # noqa: E402

F16, DEVICE = _capi.F16, _capi.DEVICE


def run(K, P, iters=30, warm=6):
    engs = [E.Engine(0) for _ in range(K)]
    for e in engs:
        e.load_hrnet(None, 32, 256, 192, seed=0)
    lib = engs[0]._lib
    rng = np.random.default_rng(0)
    ins, outs = [], []
    for e in engs:
        x = rng.standard_normal((P, 3, 256, 192)).astype(np.float16)
        d_in = e.dev_alloc(x.nbytes)
        d_out = e.dev_alloc(P * 17 * 64 * 48 * 2)
        E.check(lib.hbp_copy_h2d(e._ctx, C.c_void_p(d_in), x.ctypes.data_as(C.c_void_p), x.nbytes))
        e.sync()
        ins.append(d_in); outs.append(d_out)

    def step():
        for e, a, b in zip(engs, ins, outs):
            E.check(lib.hbp_hrnet_forward(e._ctx, C.c_void_p(a), P, C.c_void_p(b), F16, DEVICE))

    for _ in range(warm):
        step()
    for e in engs:
        e.sync()
    t0 = time.perf_counter()
    for _ in range(iters):
        step()
    for e in engs:
        e.sync()
    dt = (time.perf_counter() - t0) / iters
    print(f"contexts={K} crops_each={P} total={K*P}: {dt*1e3:.3f} ms per round, {K*P/dt:.0f} crops/s", flush=True)
    for e in engs:
        e.close()


if __name__ == "__main__":
    cases = ((1, 64), (2, 32), (4, 16), (1, 128), (2, 64), (1, 32))
    if len(sys.argv) > 2:
        cases = ((int(sys.argv[1]), int(sys.argv[2])),)
    for K, P in cases:
        run(K, P)
