"""GPU tier (-m gpu): the HRNet engine against the fp32 oracle AT THE BENCHMARKED BATCH SIZES.

Conv plans (halo / per-tap / grouped kernels, tile shapes, N splits, SM shares) are chosen per
activation-buffer capacity, so the plans behind every BENCH number (capacity 64, W32 256x192) and
behind BASELINE configs[3] (W48 384x288, 16 persons) are compared with oracle/hrnet_fp32.py here,
not only the small-batch plans of test_gpu_parity.py.  Tolerance (north_star): heatmaps within
1e-2 relative (max|a-b| / max|b| per map) of the fp32 network.  The oracle network is PARITY
UNPINNED against the reference (the reference ships no network source).
"""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def rel_err_maps(a, b):
    a, b = a.astype(np.float64), b.astype(np.float64)
    return np.abs(a - b).max(axis=(-1, -2)) / np.abs(b).max(axis=(-1, -2))


@pytest.fixture(scope="module")
def eng():
    from human_body_proportion_estimation_b200.engine import Engine
    e = Engine(0)
    yield e
    e.close()


def test_hrnet_w32_batch64_vs_fp32_oracle(eng):
    """the capacity-64 plan that produces every bench number"""
    from oracle.hrnet_fp32 import HRNetFP32
    w = eng.load_hrnet(None, 32, 256, 192, seed=0)
    crops = np.random.default_rng(64).uniform(0, 1, (64, 3, 256, 192)).astype(np.float16)
    hm = eng.hrnet_forward(crops, np.float32)
    assert np.array_equal(eng.hrnet_forward(crops, np.float32), hm)          # eager pass == graph capture
    assert np.array_equal(eng.hrnet_forward(crops, np.float32), hm)          # == graph replay
    ref = HRNetFP32(w, 32)(crops.astype(np.float32)).numpy()
    err = rel_err_maps(hm, ref)
    print("W32 batch 64: heatmap rel err max %.3e  mean %.3e" % (err.max(), err.mean()))
    assert np.isfinite(hm).all()
    assert err.max() < 1e-2
    # argmax agreement where the fp32 maximum is clear of the runner-up by more than the tolerance
    flat_r, flat_g = ref.reshape(64, 17, -1), hm.reshape(64, 17, -1)
    top2 = np.sort(flat_r, -1)[..., -2:]
    clear = (top2[..., 1] - top2[..., 0]) > 2e-2 * np.abs(flat_r).max(-1)
    assert clear.sum() > 0
    assert np.array_equal(flat_g.argmax(-1)[clear], flat_r.argmax(-1)[clear])
    # a crop's heatmaps do not depend on the batch it is in (plans differ per capacity, the K order does not)
    assert np.array_equal(eng.hrnet_forward(crops[:3], np.float32), hm[:3])


def test_hrnet_w48_batch16_vs_fp32_oracle(eng):
    """BASELINE configs[3]: HRNet-W48 384x288 on the persons of a 16-frame batch (16 persons)"""
    from oracle.hrnet_fp32 import HRNetFP32
    w = eng.load_hrnet(None, 48, 384, 288, seed=1)
    crops = np.random.default_rng(48).uniform(0, 1, (16, 3, 384, 288)).astype(np.float16)
    hm = eng.hrnet_forward(crops, np.float32)
    assert np.array_equal(eng.hrnet_forward(crops, np.float32), hm)
    ref = HRNetFP32(w, 48)(crops.astype(np.float32)).numpy()
    err = rel_err_maps(hm, ref)
    print("W48 batch 16: heatmap rel err max %.3e  mean %.3e" % (err.max(), err.mean()))
    assert hm.shape == (16, 17, 96, 72) and np.isfinite(hm).all()
    assert err.max() < 1e-2


def _stage_ops(names):
    """(oracle stage key -> program op whose output tensor it is, op index to stop after)"""
    idx = {n: i for i, n in enumerate(names)}
    out = [("conv2", idx["conv2"], idx["conv2"]), ("layer1", idx["layer1.3.conv3"], idx["layer1.3.conv3"]),
           ("transition1.0", idx["transition1.0.0"], idx["transition1.1.0.0"]),
           ("transition1.1", idx["transition1.1.0.0"], idx["transition1.1.0.0"])]
    mods = [("stage2.0", 2, 2)] + [("stage3.%d" % m, 3, 3) for m in range(4)] + \
           [("stage4.%d" % m, 4, 4 if m < 2 else 1) for m in range(3)]
    for pre, nb, n_out in mods:
        last = max(i for i, n in enumerate(names) if n.startswith(pre + "."))
        for i in range(n_out):
            if i < nb - 1:
                op = idx["%s.fuse_layers.%d.upadd" % (pre, i)]
            else:       # the lowest-resolution output is closed by the last link of its longest stride-2 chain
                op = idx["%s.fuse_layers.%d.0.%d.0" % (pre, nb - 1, nb - 2)]
            out.append(("%s.%d" % (pre, i), op, last))
    return out


def test_hrnet_w32_per_stage_error_table(eng):
    """Where the final heatmap error comes from: every stage boundary of the engine (hbp_hrnet_forward_until +
    hbp_hrnet_debug_tensor, the graph path's plans and kernels run eagerly) against the fp32 oracle's activation
    at the same boundary.  The engine stores fp16 activations and sums in fp32; no boundary may be further than
    1e-2 of the tensor's range from the fp32 network, and the error must not jump at any single stage."""
    from oracle.hrnet_fp32 import HRNetFP32
    w = eng.load_hrnet(None, 32, 256, 192, seed=0)
    P = 4
    crops = np.random.default_rng(7).uniform(0, 1, (P, 3, 256, 192)).astype(np.float16)
    stages = {}
    ref_hm = HRNetFP32(w, 32).forward(crops.astype(np.float32), stages=stages).numpy()
    names = eng.hrnet_op_names()
    rows = []
    for key, op, stop in _stage_ops(names):
        eng.hrnet_forward_until(crops, stop)
        t = eng.hrnet_debug_tensor(op).astype(np.float32)            # (P,h,w,c) NHWC
        r = np.transpose(stages[key], (0, 2, 3, 1))                   # NCHW -> NHWC
        t = t[..., :r.shape[-1]]                                      # stored channels may be padded (W48)
        assert t.shape == r.shape, (key, t.shape, r.shape)
        scale = np.abs(r).max()
        e_max = np.abs(t - r).max() / scale
        e_rms = np.sqrt(np.mean((t - r) ** 2)) / np.sqrt(np.mean(r ** 2))
        rows.append((key, e_max, e_rms))
    hm = eng.hrnet_forward(crops, np.float32)
    rows.append(("final_layer", rel_err_maps(hm, ref_hm).max(), np.sqrt(np.mean((hm - ref_hm) ** 2)) / np.sqrt(np.mean(ref_hm ** 2))))
    print("\nstage boundary                 max|d|/max|ref|   rms(d)/rms(ref)")
    for key, e_max, e_rms in rows:
        print("%-30s %12.3e %16.3e" % (key, e_max, e_rms))
    assert max(r[1] for r in rows) < 1e-2
    # fp16 storage alone gives ~5e-4 per rounding; nothing may add more than 4e-3 of range in one stage
    prev = 0.0
    for key, e_max, _ in rows:
        assert e_max - prev < 4e-3, (key, e_max, prev)
        prev = max(prev, e_max)


def test_graph_cache_alternating_batch_sizes(eng):
    """a stream whose person count changes every frame keeps one CUDA graph per batch size (and one group table
    per batch size behind it): replays stay bit-identical to the eager first pass when the sizes alternate."""
    eng.load_hrnet(None, 32, 256, 192, seed=0)
    crops = np.random.default_rng(21).uniform(0, 1, (16, 3, 256, 192)).astype(np.float16)
    big = eng.hrnet_forward(crops)                      # capacity 16 from the start: the plans do not change below
    first = {}
    for n in (5, 8, 3, 16):
        first[n] = eng.hrnet_forward(crops[:n])          # eager
        assert np.array_equal(first[n], big[:n])
    launches0 = eng.kernel_launches()
    for rep in range(3):                                 # capture on the second visit, replay afterwards
        for n in (5, 8, 3, 16, 8, 5):
            assert np.array_equal(eng.hrnet_forward(crops[:n]), first[n]), (rep, n)
    assert eng.kernel_launches() > launches0
