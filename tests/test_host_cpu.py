"""CPU tier: the C-ABI library loads and exports what include/hbp.h declares (no
compute without a GPU), host-side logic (program builder, weight packing, frame
sharding, crop geometry), and loud failure without a device."""
import os
import re

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def lib():
    from human_body_proportion_estimation_b200 import build, _capi
    build.build()
    return _capi.lib()


def test_header_symbols_exported(lib):
    from human_body_proportion_estimation_b200 import _capi
    hdr = open(os.path.join(ROOT, "include", "hbp.h")).read()
    declared = set(re.findall(r"HBP_API\s+[\w\s\*]+?\b(hbp_\w+)\s*\(", hdr))
    assert len(declared) >= 30
    for name in declared:
        assert hasattr(lib, name), name
    assert declared == set(_capi.EXPORTS)         # the ctypes table binds exactly the header
    assert lib.hbp_version() == 100


def test_no_gpu_fails_loudly(lib):
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    from human_body_proportion_estimation_b200 import _capi
    from human_body_proportion_estimation_b200.engine import Engine
    with pytest.raises(_capi.HbpError, match="no CPU fallback"):
        Engine(0)


def test_product_never_imports_oracle():
    pkg = os.path.join(ROOT, "human_body_proportion_estimation_b200")
    for f in os.listdir(pkg):
        if f.endswith(".py"):
            src = open(os.path.join(pkg, f)).read()
            assert not re.search(r"^\s*(from|import)\s+oracle\b", src, re.M), f


def test_hrnet_program(lib):
    from human_body_proportion_estimation_b200 import hrnet_arch
    rows, nw, nb = hrnet_arch.layer_table(32, 256, 192)
    assert len(rows) == 293 and nw == 28481760        # 28.48 M conv parameters (paper: 28.5 M)
    rows48, nw48, _ = hrnet_arch.layer_table(48, 384, 288)
    assert len(rows48) == 293 and nw48 == 63516144    # 63.5 M (paper: 63.6 M)
    names = [r[0] for r in rows]
    assert len(set(names)) == len(names)
    for key in ("conv1", "layer1.0.downsample.0", "transition1.1.0.0", "stage2.0.branches.1.3.conv2",
                "stage3.3.fuse_layers.2.0.1.0", "stage4.2.fuse_layers.0.3.0", "transition3.3.0.0", "final_layer"):
        assert key in names, key
    assert not any(n.startswith("stage4.2.fuse_layers.1") for n in names)   # last module fuses to branch 0 only
    # offsets are contiguous in table order
    off = 0
    for name, cin, cout, k, s, w_off, b_off in rows:
        assert w_off == off
        off += cin * cout * k * k
    w = hrnet_arch.random_weights(32, 256, 192, seed=0)
    wb, bb = hrnet_arch.pack(w)
    assert wb.dtype == np.float16 and wb.size == nw and bb.size == nb
    name, cin, cout, k, s, w_off, _ = rows[1]
    assert np.array_equal(wb[w_off:w_off + 64], w[name][0][0, :, 0, 0].astype(np.float16))   # [tap][cout][cin]


def test_oracle_hrnet_matches_independent_torch_modules(lib):
    """the functional fp32 oracle equals an nn.Module-free recomputation of one block"""
    import torch
    import torch.nn.functional as F
    from human_body_proportion_estimation_b200 import hrnet_arch
    from oracle.hrnet_fp32 import HRNetFP32
    w = hrnet_arch.random_weights(32, 64, 64, seed=3)
    net = HRNetFP32(w, 32)
    x = torch.rand(1, 3, 64, 64)
    hm = net(x)
    assert hm.shape == (1, 17, 16, 16) and torch.isfinite(hm).all()
    t = F.relu(F.conv2d(x, torch.from_numpy(w["conv1"][0]), torch.from_numpy(w["conv1"][1]), stride=2, padding=1))
    assert torch.allclose(net.cv("conv1", x, stride=2), t)


def test_frame_sharding_and_geometry():
    from human_body_proportion_estimation_b200.engine import MultiGpuEngine, lengths_to_dict
    from human_body_proportion_estimation_b200 import geometry
    plan = MultiGpuEngine.shard(10, 4)
    assert plan == [[0, 4, 8], [1, 5, 9], [2, 6], [3, 7]]
    assert sorted(sum(plan, [])) == list(range(10))
    from oracle import imgproc
    box = np.array([0.1, 0.2, 0.8, 0.5], np.float32)
    assert np.array_equal(geometry.crop_and_resize_matrices(box, 1080, 1920, 384, 288)[0],
                          imgproc.crop_and_resize_matrix(box, 1080, 1920, 384, 288))
    assert np.array_equal(geometry.box_resize_matrices([10, 20, 110, 220], 256, 192)[0],
                          imgproc.box_resize_matrix([10, 20, 110, 220], 256, 192))
    d = lengths_to_dict(np.array([1.5, 0, 0, 2, 0, 0, 0, 0, 0, 0, 0], np.float32), 3.25)
    assert d["shoulder"] == np.float32(1.5) and isinstance(d["torso"], np.float64)
    assert d["lankle_lknee"] == "Part not visible"


def test_length_helper_strict_mode_raises_like_the_reference():
    """the values themselves come from the GPU (hbp_keypoint_lengths; tests/test_gpu_entrypoints.py)"""
    from human_body_proportion_estimation_b200.pose_estimator import PoseEstimator
    k = np.zeros((17, 2), np.float32)
    with pytest.raises(UnboundLocalError):
        PoseEstimator.get_keypoint_dist_dict(0.41, k, {5}, strict=True)


def test_crop_fp16_finalize_is_exact_for_every_accumulator():
    """csrc/crop_warp.cu finish_acc<__half>: the bilinear blend is an integer sum acc = sum(byte * w), w in
    1/1024 units (0 <= acc <= 255*1024).  The fp16 output is produced with ONE fp32 multiply by RN(1/261120)
    instead of the correctly rounded fp32 quotient (acc/1024)/255 that cv2 + the reference's `/255` give;
    exhaustively: both round to the same fp16 for every possible accumulator."""
    acc = np.arange(0, 255 * 1024 + 1, dtype=np.int64)
    ref32 = (acc.astype(np.float32) / np.float32(1024.0)) / np.float32(255.0)
    y = np.float32(1.0) / np.float32(261120.0)
    one_mul = (acc.astype(np.float64) * np.float64(y)).astype(np.float32)      # RN of the exact product = the kernel's FMA
    assert np.array_equal(one_mul.astype(np.float16), ref32.astype(np.float16))
    # and the fp32 path's Markstein step reproduces the correctly rounded quotient
    q = acc.astype(np.float32) * (np.float32(1.0) / np.float32(255.0))
    r = (acc.astype(np.float64) - q.astype(np.float64) * 255.0).astype(np.float32)
    q2 = (q.astype(np.float64) + r.astype(np.float64) * np.float64(np.float32(1.0) / np.float32(255.0))).astype(np.float32)
    assert np.array_equal(q2 * np.float32(2.0 ** -10), ref32)


def test_oracle_hrnet_table_matches_library(lib):
    """the reference arm (bench.py --impl reference) builds its network from oracle/hrnet_table.py and never loads the
    library under test: both tables, both FLOP counts and both seeded weight sets must be the same"""
    from human_body_proportion_estimation_b200 import _capi, hrnet_arch
    from oracle import hrnet_table
    for width, (h, w) in ((32, (256, 192)), (48, (384, 288))):
        rows, _, _ = _capi.describe_hrnet(width, h, w)
        assert [tuple(r[:5]) for r in rows] == hrnet_table.layer_table(width)
        assert hrnet_table.flops_per_crop(width, h, w) == hrnet_arch.flops_per_crop(width, h, w)[0]
    wa, wb = hrnet_arch.random_weights(32, 256, 192, 0), hrnet_table.random_weights(32, 0)
    assert wa.keys() == wb.keys()
    for k in wa:
        assert np.array_equal(wa[k][0], wb[k][0]) and np.array_equal(wa[k][1], wb[k][1]), k


def test_reference_arm_does_not_load_the_library():
    """bench.py --impl reference in a fresh interpreter: synthetic inputs, oracle network table, oracle stages -- and no
    libhbp_b200.so in the process's memory map afterwards"""
    import subprocess
    import sys
    code = (
        "import sys; sys.path.insert(0, %r)\n"
        "import bench\n"
        "from oracle import hrnet_table, detect, geometry, imgproc, hrnet_fp32\n"
        "cfg = bench.CONFIGS[2]\n"
        "frames, dets = bench.synth_inputs(cfg)\n"
        "w = hrnet_table.random_weights(32, 0)\n"
        "d = detect.official_nms(dets[0], 0.4, 0.5, classes=[0])[0]\n"
        "assert d.shape[1] == 6 and 20 <= d.shape[0] <= 48\n"
        "print('mapped' if any('libhbp' in l for l in open('/proc/self/maps')) else 'clean')\n" % ROOT)
    out = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, timeout=300)
    assert out.returncode == 0, out.stderr[-2000:]
    assert out.stdout.strip().endswith("clean")


def test_det_pose_stream_order_and_depth_with_fake_engines():
    """MultiGpuEngine.det_pose_stream: step s goes to engine s mod G with per-engine step s // G, at most depth tickets are
    outstanding per engine, results come back in step order (host logic only: the engines are stand-ins)."""
    from human_body_proportion_estimation_b200.engine import MultiGpuEngine

    class Fake:
        def __init__(self, rank):
            self.rank, self.out, self.max_out = rank, 0, 0

        def submit(self, step):
            self.out += 1
            self.max_out = max(self.max_out, self.out)
            return (self.rank, step)

        def det_pose_collect(self, ticket, **kw):
            self.out -= 1
            return {"n": 1, "ticket": ticket, "kw": kw}

    for G, depth, n in ((2, 2, 11), (3, 1, 7), (1, 2, 5)):
        engs = [Fake(r) for r in range(G)]
        pool = MultiGpuEngine(engines=engs)
        seen = []
        res = pool.det_pose_stream(lambda e, r, s: (seen.append((r, s)), e.submit(s))[1], n, depth=depth,
                                   collect_kw={"return_heatmaps": False})
        assert [r["ticket"] for r in res] == [(s % G, s // G) for s in range(n)]
        assert seen == [(s % G, s // G) for s in range(n)]
        assert all(e.out == 0 and e.max_out <= depth for e in engs)
        assert res[0]["kw"] == {"return_heatmaps": False}
