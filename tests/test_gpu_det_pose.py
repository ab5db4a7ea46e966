"""GPU tier (-m gpu): the chained det -> pose pipeline (hbp_det_pose_submit / _collect, BASELINE configs[2] and [3])
against (a) the same stages called one by one through the C ABI with host round trips in between and (b) the oracle.

Reference chain: obj_det_yolov5_onnx.py:107-122 (letterbox -> network -> non_max_suppression) ->
modules/onnx_utils.py:252-266 (scale_coords) -> modules/pose_estimator.py:29-45 (crop = cv2.resize of the box) ->
HRNet -> person_det_pose_edet4_trtserver.py:145-171 (decode, remap, lengths); and models/conv.py:22-80 ->
person_det_pose_edet4_trtserver.py:145-171 for the EfficientDet ensemble.  Bit-exact on every output.
"""
import numpy as np
import pytest

from human_body_proportion_estimation_b200 import geometry, synth

pytestmark = pytest.mark.gpu
H, W = 1080, 1920
KEYS = ("kpts_img", "scores", "ignored", "lengths_cm", "torso_cm")


@pytest.fixture(scope="module")
def eng():
    from human_body_proportion_estimation_b200.engine import Engine
    e = Engine(0)
    yield e
    e.close()


def yolo_stage_by_stage(eng, frame, pred, heights, out_hw, conf=0.4, iou=0.5):
    """the chain's stages as separate C-ABI calls (every intermediate visits the host)"""
    from oracle import detect as od
    from oracle import imgproc
    det = eng.yolo_nms(pred, conf, iou, classes=[0])[0]
    want_det = od.official_nms(pred, conf, iou, classes=[0])[0]
    assert np.array_equal(det, want_det)                                     # NMS rows bit-exact vs the oracle
    boxes = det[:, :4].copy()
    eng.scale_coords((640, 640), boxes, frame.shape[:2])
    assert np.array_equal(boxes, od.scale_coords((640, 640), det[:, :4].copy(), frame.shape[:2]))
    ints = np.array([[int(v) for v in b] for b in boxes], np.int64).reshape(-1, 4)       # x1,y1,x2,y2 (int() truncation)
    mats = geometry.box_resize_matrices(ints, out_hw[0], out_hw[1])
    for i in range(len(ints)):
        assert np.array_equal(mats[i], imgproc.box_resize_matrix(ints[i], out_hw[0], out_hw[1]))
    boxes_yxyx = ints[:, [1, 0, 3, 2]].astype(np.float32)
    n = len(ints)
    hts = [heights[min(i, len(heights) - 1)] for i in range(n)]
    out = eng.pose_pipeline(frame, mats, np.zeros(n, np.int32), boxes_yxyx, hts, swap_rb=False, return_heatmaps=np.float16)
    return out, boxes_yxyx


def test_chain_yolo_config2(eng):
    """configs[2]: one 1080p frame, synthetic decoded YOLOv5s head (30 planted persons + 300 distractors), HRNet-W32"""
    from oracle import geometry as og
    eng.load_hrnet(None, 32, 256, 192, seed=0)
    frame = synth.frame_u8(H, W, seed=synth.SEED_BASE + 3)
    pred, _ = synth.yolo_decoded_head()
    heights = [180.0, 165.0, 172.5]
    want, boxes = yolo_stage_by_stage(eng, frame, pred, heights, (256, 192))
    n = boxes.shape[0]
    assert 20 <= n <= 40
    tk = eng.det_pose_submit_yolo(frame, pred, person_height=heights, persons_cap=48)
    got = eng.det_pose_collect(tk, return_heatmaps=True)
    assert got["n"] == n and got["status"] == 0
    assert np.array_equal(got["frame_idx"], np.zeros(n, np.int32))
    assert np.array_equal(got["boxes_yxyx_px"], boxes)
    assert np.array_equal(got["heatmaps"], want["heatmaps"])
    for k in KEYS:
        assert np.array_equal(got[k], want[k], equal_nan=True), k
    # decode + geometry of the chain against the oracle on the chain's own heatmaps
    for i in range(n):
        r = og.person_postprocess(got["heatmaps"][i].astype(np.float32), boxes[i], heights[min(i, 2)])
        assert np.array_equal(got["kpts_img"][i], r["xy_img"])
        wl = og.lengths_to_array(r["lengths"])
        gl = got["lengths_cm"][i].astype(np.float64)
        gl[1] = got["torso_cm"][i]
        assert np.array_equal(gl, wl)
    # two batches in flight, a frame without detections in between, the PIL-bicubic letterbox variant
    empty = pred.copy()
    empty[..., 4] = 0.0
    t0 = eng.det_pose_submit_yolo(frame, pred, person_height=heights, persons_cap=48, resample="bicubic")
    t1 = eng.det_pose_submit_yolo(frame, empty, person_height=heights, persons_cap=48)
    g0, g1 = eng.det_pose_collect(t0), eng.det_pose_collect(t1)
    assert g1["n"] == 0 and g1["kpts_img"].shape == (0, 17, 2)
    for k in KEYS:
        assert np.array_equal(g0[k], want[k], equal_nan=True), k
    # capacity overflow is reported, the first persons_cap persons are still right
    tk = eng.det_pose_submit_yolo(frame, pred, person_height=heights, persons_cap=8)
    g = eng.det_pose_collect(tk)
    assert g["n"] == 8 and (g["status"] & 2)
    for k in KEYS:
        assert np.array_equal(g[k], want[k][:8], equal_nan=True), k
    tk = eng.det_pose_submit_yolo(frame, pred, person_height=heights, persons_cap=48, cand_cap=64)
    assert eng.det_pose_collect(tk)["status"] & 1


def test_det_pose_stream_two_contexts_on_one_gpu(eng):
    """MultiGpuEngine(devices=[0, 0]).det_pose_stream: steps alternate between two engine contexts on the same GPU (their
    forwards overlap); every step's outputs equal the single-engine call on the same inputs, in step order."""
    from human_body_proportion_estimation_b200.engine import MultiGpuEngine
    eng.load_hrnet(None, 32, 256, 192, seed=0)
    pred, _ = synth.yolo_decoded_head()
    frames = [synth.frame_u8(H, W, seed=synth.SEED_BASE + 30 + i) for i in range(3)]
    heights = [180.0, 165.0, 172.5]
    want = [eng.det_pose_collect(eng.det_pose_submit_yolo(f, pred, person_height=heights, persons_cap=48)) for f in frames]
    import ctypes as C
    from human_body_proportion_estimation_b200 import _capi
    n_dev = C.c_int()
    _capi.check(_capi.lib().hbp_device_count(C.byref(n_dev)))
    devices = [d for d in range(min(n_dev.value, 2)) for _ in range(2)]        # two contexts on each of (up to) two GPUs
    pool = MultiGpuEngine(devices=devices, width=32, in_h=256, in_w=192, seed=0)
    G = len(devices)
    n_steps = 4 * G + 1
    got = pool.det_pose_stream(lambda e, r, s: e.det_pose_submit_yolo(frames[(s * G + r) % 3], pred, person_height=heights, persons_cap=48),
                               n_steps, depth=2)
    assert len(got) == n_steps
    for step, g in enumerate(got):
        w = want[step % 3]                      # step -> engine step % G, per-engine step step // G: frame (s*G + r) % 3 = step % 3
        assert g["n"] == w["n"] and g["status"] == 0
        for k in KEYS + ("boxes_yxyx_px",):
            assert np.array_equal(g[k], w[k], equal_nan=True), (step, k)
    for e in pool.engines:
        e.close()


def test_chain_edet_config3(eng):
    """configs[3]: 16 frames 1080p, synthetic EfficientDet outputs, up to 16 persons per frame, HRNet-W48 384x288"""
    from oracle import detect as od
    eng.load_hrnet(None, 48, 384, 288, seed=1)
    F, per = 4, 5                                   # (the full 16 x 16 shape runs in bench.py --config 3)
    frames = np.stack([synth.frame_u8(H, W, seed=synth.SEED_BASE + 4 + i) for i in range(F)])
    boxes, scores, classes = synth.edet_outputs(F, per, H, W)
    heights = [175.0, 160.0]
    mats, bpx, fidx, hts = [], [], [], []
    for f in range(F):
        bn = eng.edet_person_filter(boxes[f], scores[f], classes[f], 0.70, H // 17, 0, H, W, max_persons=8)[0]
        want_bn, _ = od.edet_person_filter(boxes[f], scores[f], classes[f], 0.70, H // 17, 0, H, W, 8)
        assert np.array_equal(bn, want_bn)
        mats.append(geometry.crop_and_resize_matrices(bn, H, W, 384, 288))
        b = bn.copy()
        b *= [H, W, H, W]                           # person_det_pose_edet4_trtserver.py:145
        bpx.append(b)
        fidx += [f] * len(bn)
        hts += [heights[min(i, 1)] for i in range(len(bn))]
    mats, bpx = np.concatenate(mats), np.concatenate(bpx)
    n = len(fidx)
    assert n == F * per
    want = eng.pose_pipeline(frames, mats, np.asarray(fidx, np.int32), bpx, hts, swap_rb=False)
    tk = eng.det_pose_submit_edet(frames, boxes, scores, classes, person_height=heights, max_persons=8)
    got = eng.det_pose_collect(tk)
    assert got["n"] == n and got["status"] == 0
    assert np.array_equal(got["frame_idx"], np.asarray(fidx, np.int32))
    assert np.array_equal(got["boxes_yxyx_px"], bpx)
    for k in KEYS:
        assert np.array_equal(got[k], want[k], equal_nan=True), k
    # reference default: at most 3 persons per frame (models/conv.py:34-35)
    tk = eng.det_pose_submit_edet(frames, boxes, scores, classes, person_height=heights)
    g3 = eng.det_pose_collect(tk)
    assert g3["n"] == 3 * F
    sel = np.concatenate([np.arange(f * per, f * per + 3) for f in range(F)])
    for k in ("kpts_img", "scores", "ignored"):
        assert np.array_equal(g3[k], want[k][sel], equal_nan=True), k


def test_decode_inverse_affine(eng):
    """optional (P,6) dst->src matrices in K6: keypoints of rotated / aspect-padded crops map back through the crop's own
    affine (north_star item 5); the default stays the reference's box formula."""
    hm = synth.heatmaps(6, 17, 64, 48, seed=55)
    boxes = synth.person_boxes_yxyx_px(6, seed=56)
    rng = np.random.default_rng(57)
    mats = np.zeros((6, 2, 3))
    for i in range(6):
        ang, sc = rng.uniform(-0.6, 0.6), rng.uniform(0.8, 3.0)
        mats[i] = [[sc * np.cos(ang), -sc * np.sin(ang), rng.uniform(0, 900)], [sc * np.sin(ang), sc * np.cos(ang), rng.uniform(0, 500)]]
    out = eng.decode_proportions(hm, boxes, 175.0, mats=mats, crop_hw=(256, 192))
    base = eng.decode_proportions(hm, boxes, 175.0)
    assert np.array_equal(out["kpts_hm"], base["kpts_hm"]) and np.array_equal(out["scores"], base["scores"])
    # oracle: (u,v) = (x*crop_w/Wh, y*crop_h/Hh), image = M (u,v,1) in float64, one rounding to float32
    u = out["kpts_hm"][..., 0].astype(np.float64) * 192 / 48
    v = out["kpts_hm"][..., 1].astype(np.float64) * 256 / 64
    wx = (mats[:, 0, 0, None] * u + mats[:, 0, 1, None] * v) + mats[:, 0, 2, None]
    wy = (mats[:, 1, 0, None] * u + mats[:, 1, 1, None] * v) + mats[:, 1, 2, None]
    assert np.array_equal(out["kpts_img"][..., 0], wx.astype(np.float32))
    assert np.array_equal(out["kpts_img"][..., 1], wy.astype(np.float32))
    # lengths follow the remapped keypoints: same as the oracle's segment code on them
    from oracle import geometry as og
    for i in range(6):
        ign = {j for j in range(17) if (out["ignored"][i] >> j) & 1}
        y1, y2 = int(boxes[i, 0]), int(boxes[i, 2])
        d = og.lengths_dict(175.0 / (y2 - y1), out["kpts_img"][i], ign)
        gl = out["lengths_cm"][i].astype(np.float64)
        gl[1] = out["torso_cm"][i]
        assert np.array_equal(gl, og.lengths_to_array(d))
    # an axis-aligned crop_and_resize matrix reproduces the crop sampling grid exactly: heatmap cell (x,y) -> the source
    # position hbp_crop_warp sampled for crop pixel (4x, 4y)
    bn = boxes / np.array([1080, 1920, 1080, 1920], np.float32)
    m2 = geometry.crop_and_resize_matrices(bn, 1080, 1920, 256, 192)
    o2 = eng.decode_proportions(hm, boxes, 175.0, mats=m2, crop_hw=(256, 192))
    assert np.allclose(o2["kpts_img"][..., 0], m2[:, 0, 2, None] + m2[:, 0, 0, None] * u, rtol=0, atol=1e-3)
