"""GPU tier (-m gpu): every stage through the C ABI against the oracle and the
golden vectors the reference produced.  Bit-exact for index/integer/u8 work and
for every float path whose reference arithmetic is a fixed sequence of IEEE
operations; stated tolerances elsewhere."""
import hashlib

import numpy as np
import pytest

from human_body_proportion_estimation_b200 import synth

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def eng():
    from human_body_proportion_estimation_b200.engine import Engine
    e = Engine(0)
    yield e
    e.close()


def oracle_loop(hm, boxes, heights):
    from oracle import geometry
    n = hm.shape[0]
    o = dict(kpts_hm=np.zeros((n, 17, 2), np.float32), scores=np.zeros((n, 17), np.float32),
             kpts_img=np.zeros((n, 17, 2), np.float32), ignored=np.zeros(n, np.uint32),
             lengths=np.zeros((n, 11)), argmax=np.zeros((n, 17), np.int64))
    for i in range(n):
        r = geometry.person_postprocess(hm[i].astype(np.float32), boxes[i], heights[i])
        o["kpts_hm"][i], o["scores"][i], o["kpts_img"][i] = r["xy_hm"], r["score"][:, 0], r["xy_img"]
        o["ignored"][i] = sum(1 << j for j in r["ignored"])
        o["lengths"][i] = geometry.lengths_to_array(r["lengths"])
        o["argmax"][i] = r["idx"]
    return o


def check_decode(eng, hm, boxes, heights, g=None):
    out = eng.decode_proportions(hm, boxes, heights)
    o = oracle_loop(hm, boxes, heights)
    assert np.array_equal(out["argmax"], o["argmax"])                       # bit-exact indices
    assert np.array_equal(out["kpts_hm"], o["kpts_hm"])
    assert np.array_equal(out["scores"], o["scores"], equal_nan=True)
    assert np.array_equal(out["kpts_img"], o["kpts_img"])
    assert np.array_equal(out["ignored"], o["ignored"])
    got = out["lengths_cm"].astype(np.float64)
    got[:, 1] = out["torso_cm"]
    assert np.array_equal(got, o["lengths"])                                # incl. float64 torso
    if g is not None:      # and against what the reference itself returned
        ok = ~g["raised"]
        assert np.array_equal(out["kpts_hm"], g["xy_hm"])
        assert np.array_equal(out["kpts_img"], g["xy_img"])
        assert np.array_equal(out["ignored"], g["ignored"])
        assert np.array_equal(got[ok], g["lengths"][ok])
    return out


def test_decode_small_edge_cases(eng, golden):
    g = golden("decode_small.npz")
    hs = [float(v) for v in g["p_height"]]
    heights = [hs[min(i, len(hs) - 1)] for i in range(g["heatmaps"].shape[0])]
    out = check_decode(eng, g["heatmaps"], g["boxes_px"], heights, g)
    # the reference raises for person 3 (shoulder lost); we return "not visible" there
    assert g["raised"][3]
    assert out["lengths_cm"][3, 0] == 0 and out["lengths_cm"][3, 1] == 0


def test_decode_config1(eng, golden):
    g = golden("decode_cfg1.npz")
    check_decode(eng, synth.heatmaps(32), synth.person_boxes_yxyx_px(32), [175] * 32, g)
    g = golden("decode_cfg1_drop.npz")
    check_decode(eng, synth.heatmaps(32, seed=synth.SEED_BASE + 101, keep_torso=False),
                 synth.person_boxes_yxyx_px(32), [175] * 32, g)


def test_decode_96x72_and_fp16(eng, golden):
    g = golden("decode_96x72.npz")
    hm = synth.heatmaps(6, 17, 96, 72, seed=77)
    boxes = synth.person_boxes_yxyx_px(6, seed=78)
    check_decode(eng, hm, boxes, [180, 165, 165, 165, 165, 165], g)
    h16 = hm.astype(np.float16)
    check_decode(eng, h16, boxes, [180.0] * 6)            # fp16 maps: oracle on the same values
    # decode-only call, odd sizes (unaligned rows), J != 17
    odd = synth.heatmaps(5, 9, 13, 7, seed=3)
    out = eng.decode_proportions(odd)
    from oracle import geometry
    for i in range(5):
        xy, sc, idx = geometry.decode_heatmap(odd[i])
        assert np.array_equal(out["kpts_hm"][i], xy) and np.array_equal(out["argmax"][i], idx)


def test_decode_properties_large(eng):
    """size-independent properties at a batched size (2048 persons): the decoded
    index really is the first maximum; planting a larger value moves it there."""
    hm = synth.heatmaps(2048, seed=9)
    out = eng.decode_proportions(hm)
    flat = hm.reshape(2048, 17, -1)
    assert np.array_equal(out["argmax"], flat.argmax(-1))
    assert np.array_equal(out["scores"], flat.max(-1))
    hm[:, :, 63, 47] = 5.0
    out = eng.decode_proportions(hm)
    assert (out["argmax"] == 64 * 48 - 1).all()
    q = eng.decode_proportions(synth.heatmaps(8, seed=10), quarter_offset=True)
    from oracle import geometry
    h8 = synth.heatmaps(8, seed=10)
    for i in range(8):
        xy, _, _ = geometry.decode_heatmap(h8[i])
        assert np.array_equal(q["kpts_hm"][i], geometry.quarter_offset(h8[i], xy))


def test_crop_vs_cv2_golden(eng, golden):
    g = golden("crop_small.npz")
    img, mats = g["img"], g["mats"]
    fi = np.zeros(len(mats), np.int32)
    f32 = eng.crop_warp(img, mats, fi, 64, 48, swap_rb=False, out_dtype=np.float32)
    want = np.transpose(g["warp_f32"], (0, 3, 1, 2)) / np.float32(255)
    assert np.array_equal(f32, want.astype(np.float32))                  # bit-exact cv2 fixed point
    # the north-star gate: within 1/255 of cv2.warpAffine on the u8 image
    assert np.abs(f32 * 255 - np.transpose(g["warp_u8"], (0, 3, 1, 2))).max() <= 0.5 + 1e-3
    f16 = eng.crop_warp(img, mats, fi, 64, 48, swap_rb=True, out_dtype=np.float16)
    assert np.array_equal(f16, want[:, ::-1].astype(np.float16))


def test_crop_config2_shapes(eng):
    """config 2: one 1080p frame (white noise = worst case), 64 boxes, 256x192 fp16."""
    from oracle import imgproc
    frame = synth.frame_u8(smooth=False)
    boxes = synth.person_boxes_yxyx_px(64, seed=synth.SEED_BASE + 2, hmin=150, hmax=900)
    boxes_n = boxes / np.array([1080, 1920, 1080, 1920], np.float32)
    from human_body_proportion_estimation_b200 import geometry
    mats = geometry.crop_and_resize_matrices(boxes_n, 1080, 1920, 256, 192)
    # plus rotated / out-of-frame / huge-zoom-out cases (global-memory path)
    mats[5] = [[0.9, -0.4, 300.0], [0.4, 0.9, -40.0]]
    mats[6] = [[9.5, 0.0, -30.0], [0.0, 4.1, 5.0]]
    mats[7] = [[0.1, 0.0, 1900.5], [0.0, 0.1, 1070.25]]
    got = eng.crop_warp(frame, mats, np.zeros(64, np.int32), 256, 192, swap_rb=True)
    for p in list(range(10)) + [31, 63]:
        want = imgproc.crop_persons(frame, [mats[p]], 256, 192, swap_rb=True)[0]
        assert np.array_equal(got[p], want), p
    # two frames, persons pointing at either
    frames = np.stack([frame, frame[::-1].copy()])
    fi = (np.arange(64) % 2).astype(np.int32)
    got2 = eng.crop_warp(frames, mats, fi, 256, 192)
    assert np.array_equal(got2[0], got[0])
    assert np.array_equal(got2[1], imgproc.crop_persons(frames[1], [mats[1]], 256, 192)[0])
    assert eng.crop_warp(frame, mats[:0], fi[:0], 256, 192).shape == (0, 3, 256, 192)


def test_preprocess(eng, golden):
    from human_body_proportion_estimation_b200._capi import NHWC, PRE_COPY, PRE_LETTERBOX, PRE_STRETCH
    from oracle import imgproc
    g = golden("crop_small.npz")
    img = g["img"]
    # A3: HRNet preprocess = what the reference's PoseEstimator.preprocess returned
    pre = eng.preprocess(img, PRE_STRETCH, 96, 72, True, 128, np.float32)
    assert np.array_equal(pre, g["hrnet_pre_72x96"])
    # cv2.resize u8 bit-exact, both sizes stored in the golden file
    r = eng.preprocess(img, PRE_STRETCH, 96, 72, False, 128, np.uint8, NHWC)[0]
    assert np.array_equal(r, g["resize_72x96"])
    r = eng.preprocess(img, PRE_STRETCH, 201, 333, False, 128, np.uint8, NHWC)[0]
    assert np.array_equal(r, g["resize_333x201"])
    # A1: BGR->RGB only (vectorised stream kernel) incl. a pixel count not divisible by 16
    frame = synth.frame_u8(270, 481, seed=5, smooth=False)
    c = eng.preprocess(frame, PRE_COPY, None, None, True, 128, np.uint8, NHWC)[0]
    assert np.array_equal(c, imgproc.edet_preprocess(frame))
    # A2: letterbox geometry + pad on a 1080p frame
    big = synth.frame_u8(seed=6)
    lb = eng.preprocess(big, PRE_LETTERBOX, 640, 640, False, 128, np.float32)[0]
    assert np.array_equal(lb, imgproc.letterbox_linear(big, 640, 640))
    assert (lb[:, :140] == np.float32(128) / np.float32(255)).all() and (lb[:, 500:] == lb[0, 0, 0]).all()
    lb16 = eng.preprocess(np.stack([big, big]), PRE_LETTERBOX, 640, 640, False, 128, np.float16)
    assert np.array_equal(lb16[1], lb.astype(np.float16))


def test_letterbox_pil_bicubic(eng, golden):
    """A2 with the reference's own resampler: letterbox_image (PIL BICUBIC, onnx_utils.py:225-235) executed
    by the reference -> golden; the CUDA two-pass kernel reproduces it bit for bit (u8), and the f32/f16
    CHW /255 form obj_det_yolov5_onnx.py:27-36 feeds the detector."""
    import hashlib
    from human_body_proportion_estimation_b200._capi import NHWC, PRE_LETTERBOX_PIL
    from human_body_proportion_estimation_b200 import obj_det_yolov5, onnx_utils
    from oracle import imgproc
    g = golden("letterbox_pil.npz")
    i = 0
    while "img%d" % i in g:
        w, h = (int(v) for v in g["size%d" % i])
        got = eng.preprocess(g["img%d" % i], PRE_LETTERBOX_PIL, h, w, False, 128, np.uint8, NHWC)[0]
        assert np.array_equal(got, g["lb%d" % i]), i
        i += 1
    frame = synth.frame_u8(smooth=False)                       # config 3's 1080p frame, white noise
    u8 = onnx_utils.letterbox_image(frame, (640, 640), engine=eng)
    assert np.array_equal(u8[::8, ::8], g["lb1080_sample"])
    assert hashlib.sha256(np.ascontiguousarray(u8).tobytes()).hexdigest() == str(g["lb1080_sha256"])
    chw = obj_det_yolov5.preprocess_image(frame, (640, 640), engine=eng)
    assert np.array_equal(chw, imgproc.letterbox_pil(frame, 640, 640))
    # batch of two 4K frames, fp16 NCHW with the BGR->RGB swap
    big = synth.frame_u8(2160, 3840, seed=9, smooth=False)
    two = eng.preprocess(np.stack([big, big[::-1]]), PRE_LETTERBOX_PIL, 640, 640, True, 128, np.float16)
    assert np.array_equal(two[0], imgproc.letterbox_pil(big[..., ::-1], 640, 640).astype(np.float16))
    assert np.array_equal(two[1], imgproc.letterbox_pil(big[::-1, :, ::-1], 640, 640).astype(np.float16))


def unpack(arr, cnt):
    return [None if c < 0 else arr[i, :c] for i, c in enumerate(cnt)]


def test_official_nms_golden(eng, golden):
    g = golden("nms_small.npz")
    pred = g["pred"]
    for key, kw in (("off_a", dict(conf_thres=0.4, iou_thres=0.5)),
                    ("off_b", dict(conf_thres=0.4, iou_thres=0.5, classes=[0, 3])),
                    ("off_c", dict(conf_thres=0.25, iou_thres=0.45))):
        got = eng.yolo_nms(pred, **kw)
        for w, o in zip(unpack(g[key], g[key + "_n"]), got):
            assert o.shape == w.shape and np.array_equal(o, w), key      # keep-set and rows bit-exact
    g3 = golden("nms_cfg3.npz")
    p3, _ = synth.yolo_decoded_head()
    assert np.array_equal(eng.yolo_nms(p3, 0.4, 0.5)[0], g3["all_cls"][0, :g3["all_cls_n"][0]])
    assert np.array_equal(eng.yolo_nms(p3, 0.4, 0.5, classes=[0])[0], g3["person"][0, :g3["person_n"][0]])
    # nothing above threshold -> empty
    assert eng.yolo_nms(p3 * 0, 0.4, 0.5)[0].shape == (0, 6)


def test_nms_known_answers_and_ties(eng, golden):
    """torchvision keep-sets through the full path: single-class heads whose boxes are
    integers (so xywh<->xyxy is exact) reproduce the stored keep indices."""
    from oracle import detect
    g = golden("nms_kat.npz")
    b, s = g["boxes"], g["scores"]
    pred = np.zeros((1, len(b), 6), np.float32)
    pred[0, :, 0] = (b[:, 0] + b[:, 2]) / 2
    pred[0, :, 1] = (b[:, 1] + b[:, 3]) / 2
    pred[0, :, 2] = b[:, 2] - b[:, 0]
    pred[0, :, 3] = b[:, 3] - b[:, 1]
    pred[0, :, 4] = s
    pred[0, :, 5] = 1.0
    got = eng.yolo_nms(pred, 0.1, 0.5)[0]
    assert np.array_equal(got[:, :4], b[g["keep"]]) and np.array_equal(got[:, 4], s[g["keep"]])
    # random boxes with many tied scores: compare with the oracle on the same head
    rng = np.random.default_rng(4)
    n = 3000
    pred = np.zeros((2, n, 5 + 3), np.float32)
    pred[..., 0:2] = rng.integers(20, 620, (2, n, 2))
    pred[..., 2:4] = rng.integers(4, 80, (2, n, 2)) * 2
    pred[..., 4] = np.round(rng.uniform(0.3, 1, (2, n)), 2)
    pred[..., 5:] = np.round(rng.uniform(0.5, 1, (2, n, 3)), 1)
    for thr in (0.3, 0.5, 0.7):
        want = detect.official_nms(pred, 0.35, thr)
        got = eng.yolo_nms(pred, 0.35, thr)
        for w, o in zip(want, got):
            assert np.array_equal(o, w)
    # rows whose byte span is not 16-byte aligned (E = 7 floats, N = 1001, three images): the streaming filter's unaligned
    # head / tail vectors, and a batch of 9 images for the sector-wise filter on the same data
    for B_ in (3, 9):
        pu = np.zeros((B_, 1001, 5 + 2), np.float32)
        pu[..., 0:2] = rng.integers(20, 620, (B_, 1001, 2))
        pu[..., 2:4] = rng.integers(4, 80, (B_, 1001, 2)) * 2
        pu[..., 4] = np.round(rng.uniform(0.2, 1, (B_, 1001)), 2)
        pu[..., 5:] = np.round(rng.uniform(0.3, 1, (B_, 1001, 2)), 1)
        want = detect.official_nms(pu, 0.35, 0.5)
        got = eng.yolo_nms(pu, 0.35, 0.5)
        for w, o in zip(want, got):
            assert np.array_equal(o, w), B_
    # the three regimes of the sweep kernel (csrc/detect.cu: nms_sweep_gather_kernel): mask staged in shared memory with the
    # removed words in registers (<= 1024 candidates), staged with more than 32 words, and on the mask in global memory
    for k_cand in (40, 900, 1150, 2400):
        p1 = pred[:1].copy()
        _, src = detect.yolo_candidates(p1[0], 0.35)
        assert len(src) >= k_cand
        p1[0, src[k_cand:], 4] = 0.0                      # exactly k_cand candidates left
        assert len(detect.yolo_candidates(p1[0], 0.35)[1]) == k_cand
        for max_det in (300, 7):
            want = detect.official_nms(p1, 0.35, 0.5, max_det=max_det)
            got = eng.yolo_nms(p1, 0.35, 0.5, max_det=max_det)
            assert np.array_equal(got[0], want[0]), (k_cand, max_det)


def test_legacy_nms(eng, golden):
    from oracle import detect
    g = golden("nms_small.npz")
    got = eng.yolo_nms_legacy(g["pred"], 8, 0.4, 0.3)
    want_ref = unpack(g["leg"], g["leg_n"])
    want_or = detect.legacy_nms(g["pred"], 8, 0.4, 0.3)
    for o, wr, wo in zip(got, want_ref, want_or):
        assert np.array_equal(o, wo)                       # oracle: stable tie order
        if len(np.unique(wr[:, 4])) == len(wr):
            assert np.array_equal(o, wr)                   # reference, where its unstable sort is defined
    empty = eng.yolo_nms_legacy(g["pred"] * 0, 8, 0.4, 0.3)
    assert empty[0] is None and empty[1] is None
    # wrapper mutates its input like the reference (onnx_utils.py:47)
    from human_body_proportion_estimation_b200 import onnx_utils
    p = g["pred"].copy()
    onnx_utils.w_non_max_suppression(p, 8, 0.4, 0.3, engine=eng)
    assert np.array_equal(p[..., :4], g["leg_mutated"][..., :4])


def test_yolo_raw_decode(eng):
    from oracle import detect
    rng = np.random.default_rng(0)
    heads = [rng.normal(0, 2, (2, 3, s, s, 85)).astype(np.float32) for s in (20, 40, 80)]
    got = eng.yolo_decode_raw(heads)
    want = detect.yolo_raw_decode(heads)
    assert got.shape == want.shape == (2, 25200, 85)
    # sigmoid goes through exp: tolerance 4e-6 relative (float32, a few ulp of expf)
    np.testing.assert_allclose(got, want, rtol=4e-6, atol=1e-6)


def test_edet_filter_and_scale_coords(eng, golden):
    from oracle import detect
    boxes, scores, classes = synth.edet_outputs()
    for maxp in (3, 16):
        got = eng.edet_person_filter(boxes, scores, classes, 0.7, 1080 // 17, 0, 1080, 1920, max_persons=maxp)
        for f in range(boxes.shape[0]):
            want, _ = detect.edet_person_filter(boxes[f], scores[f], classes[f], 0.7, 1080 // 17, 0, 1080, 1920, maxp)
            assert np.array_equal(got[f], want)
    none = eng.edet_person_filter(boxes, scores * 0, classes, 0.7, 0, 0, 1080, 1920)
    assert all(x.shape == (0, 4) for x in none)
    g = golden("misc.npz")
    c = g["coords"].copy()
    assert np.array_equal(eng.scale_coords((640, 640), c, (1080, 1920)), g["scaled_1080"])
    c = g["coords"].copy()
    assert np.array_equal(eng.scale_coords((640, 640), c, (2160, 3840)), g["scaled_2160"])


def rel_err(a, b):
    """max|a-b| / max|b| per heatmap (element-relative error is
    unbounded at zero crossings)."""
    a, b = a.astype(np.float64), b.astype(np.float64)
    return (np.abs(a - b).max(axis=(-1, -2)) / np.abs(b).max(axis=(-1, -2))).max()


@pytest.fixture(scope="module")
def hrnet32(eng):
    from oracle.hrnet_fp32 import HRNetFP32
    w = eng.load_hrnet(None, 32, 256, 192, seed=0)
    rng = np.random.default_rng(5)
    crops = rng.uniform(0, 1, (3, 3, 256, 192)).astype(np.float16)
    ref = HRNetFP32(w, 32)(crops.astype(np.float32)).numpy()
    return w, crops, ref


def test_hrnet_w32_simt_engine(eng, hrnet32):
    """engine 0 (SIMT tiles) against the fp32 torch model: 1e-2 of the map's range."""
    _, crops, ref = hrnet32
    eng.set_hrnet_engine(0)
    hm = eng.hrnet_forward(crops, np.float32)
    assert rel_err(hm, ref) < 1e-2
    # second and third call replay the captured CUDA graph: identical bits
    assert np.array_equal(eng.hrnet_forward(crops, np.float32), hm)
    assert np.array_equal(eng.hrnet_forward(crops, np.float32), hm)


def test_hrnet_w32_tcgen05_engine(eng, hrnet32):
    """engine 1 (tcgen05/TMEM implicit GEMM fed by TMA) against the fp32 model and
    against engine 0; batch sizes that do not fill the image tiles included."""
    _, crops, ref = hrnet32
    eng.set_hrnet_engine(1)
    hm = eng.hrnet_forward(crops, np.float32)
    assert np.isfinite(hm).all()
    assert rel_err(hm, ref) < 1e-2
    assert np.array_equal(eng.hrnet_forward(crops, np.float32), hm)
    assert np.array_equal(eng.hrnet_forward(crops, np.float32), hm)
    one = eng.hrnet_forward(crops[:1], np.float32)
    assert np.array_equal(one[0], hm[0])                # images are independent
    h16 = eng.hrnet_forward(crops, np.float16)
    assert np.array_equal(h16, hm.astype(np.float16))


def test_pipeline_end_to_end(eng, hrnet32):
    """frame + boxes -> crop -> HRNet -> decode through hbp_pose_pipeline equals the
    stage-by-stage calls, and its decode is bit-exact on its own heatmaps."""
    from human_body_proportion_estimation_b200 import geometry
    eng.set_hrnet_engine(1)
    frame = synth.frame_u8(seed=11)
    boxes = synth.person_boxes_yxyx_px(5, seed=12)
    boxes_n = boxes / np.array([1080, 1920, 1080, 1920], np.float32)
    mats = geometry.crop_and_resize_matrices(boxes_n, 1080, 1920, 256, 192)
    fi = np.zeros(5, np.int32)
    out = eng.pose_pipeline(frame, mats, fi, boxes, 175, return_heatmaps=np.float16)
    crops = eng.crop_warp(frame, mats, fi, 256, 192)
    hm = eng.hrnet_forward(crops, np.float16)
    assert np.array_equal(out["heatmaps"], hm)
    o = oracle_loop(hm, boxes, [175] * 5)
    assert np.array_equal(out["kpts_img"], o["kpts_img"])
    assert np.array_equal(out["ignored"], o["ignored"])
    lengths = out["lengths_cm"].astype(np.float64)
    lengths[:, 1] = out["torso_cm"]
    assert np.array_equal(lengths, o["lengths"])
    # against the fp32 network: heatmaps within 1e-2, keypoints within 0.5 px wherever the
    # fp32 maximum is separated from the runner-up by more than the heatmap tolerance
    from oracle.hrnet_fp32 import HRNetFP32
    ref = HRNetFP32(hrnet32[0], 32)(crops.astype(np.float32)).numpy()
    assert rel_err(hm.astype(np.float32), ref) < 1e-2
    o32 = oracle_loop(ref, boxes, [175] * 5)
    flat = ref.reshape(5, 17, -1)
    top2 = np.sort(flat, -1)[..., -2:]
    clear = (top2[..., 1] - top2[..., 0]) > 2e-2 * np.abs(flat).max(-1)
    d = np.abs(out["kpts_img"] - o32["kpts_img"]).max(-1)
    assert clear.sum() > 0 and (d[clear] <= 0.5).all()


def test_pipeline_async_matches_sync(eng, hrnet32):
    """hbp_pose_pipeline_submit/_collect (two batches in flight, frame upload overlapped with the network of the
    previous batch) returns exactly what the synchronous hbp_pose_pipeline returns, batch after batch."""
    from human_body_proportion_estimation_b200 import geometry
    frames = [synth.frame_u8(seed=40 + i) for i in range(3)]
    batches = []
    for i, fr in enumerate(frames):
        n = 5 + 3 * i
        boxes = synth.person_boxes_yxyx_px(n, seed=synth.SEED_BASE + 20 + i, hmin=200, hmax=800)
        mats = geometry.crop_and_resize_matrices(boxes / np.array([1080, 1920, 1080, 1920], np.float32), 1080, 1920, 256, 192)
        batches.append((fr, mats, np.zeros(n, np.int32), boxes))
    want = [eng.pose_pipeline(fr, mats, fi, boxes, 170 + i) for i, (fr, mats, fi, boxes) in enumerate(batches)]
    tickets, got = [], []
    for i, (fr, mats, fi, boxes) in enumerate(batches):
        tickets.append(eng.pose_pipeline_submit(fr, mats, fi, boxes, 170 + i))
        if len(tickets) == 2:                              # at most two in flight
            got.append(eng.pose_pipeline_collect(tickets.pop(0)))
    while tickets:
        got.append(eng.pose_pipeline_collect(tickets.pop(0)))
    for w, g in zip(want, got):
        for k in ("kpts_img", "scores", "ignored", "lengths_cm", "torso_cm"):
            assert np.array_equal(w[k], g[k], equal_nan=True), k
    # a third submit without a collect is refused, not queued silently
    t0 = eng.pose_pipeline_submit(*batches[0], 175)
    t1 = eng.pose_pipeline_submit(*batches[1], 175)
    with pytest.raises(Exception):
        eng.pose_pipeline_submit(*batches[2], 175)
    eng.pose_pipeline_collect(t0); eng.pose_pipeline_collect(t1)
    # a frame without persons goes through the same calls and comes back empty
    fr, mats, fi, boxes = batches[0]
    empty = eng.pose_pipeline_collect(eng.pose_pipeline_submit(fr, mats[:0], fi[:0], boxes[:0], 175))
    assert empty["kpts_img"].shape == (0, 17, 2) and empty["lengths_cm"].shape == (0, 11)


def test_multi_gpu_engine_stream_matches_sync():
    """configs[4] path: MultiGpuEngine.stream (frame f -> GPU f mod G, two frames in flight per GPU) returns per frame
    what the synchronous single-GPU call returns; runs on however many GPUs the box has (1 at least)."""
    from human_body_proportion_estimation_b200 import geometry
    from human_body_proportion_estimation_b200.engine import MultiGpuEngine
    mg = MultiGpuEngine(width=32, in_h=256, in_w=192, seed=0)
    frames = [synth.frame_u8(540, 960, seed=70 + i) for i in range(5)]
    sets = []
    for i in range(5):
        boxes = synth.person_boxes_yxyx_px(7, 540, 960, seed=synth.SEED_BASE + 80 + i, hmin=100, hmax=400)
        mats = geometry.crop_and_resize_matrices(boxes / np.array([540, 960, 540, 960], np.float32), 540, 960, 256, 192)
        sets.append((mats.reshape(-1, 6), boxes))
    res, lat = mg.stream(lambda f: (frames[f], sets[f][0], sets[f][1], 170.0 + f), 5)
    e0 = mg.engines[0]
    for f in range(5):
        want = e0.pose_pipeline(frames[f], sets[f][0], np.zeros(7, np.int32), sets[f][1], 170.0 + f)
        for k in ("kpts_img", "scores", "ignored", "lengths_cm", "torso_cm"):
            assert np.array_equal(want[k], res[f][k], equal_nan=True), (f, k)
    assert len(lat) == 5 and min(lat) > 0


def test_hrnet_outputs_independent_of_capacity():
    """Conv plans (halo vs per-tap kernel, tile shapes, N splits) are chosen per activation-buffer capacity; every
    kernel sums K in the same order (Cin chunk, dx, dy, 16-channel step), so a crop's heatmaps are bit-identical
    whatever the capacity and whatever else is in the batch."""
    from human_body_proportion_estimation_b200.engine import Engine
    e = Engine(0)
    e.load_hrnet(None, 32, 256, 192, seed=0)
    crops = np.random.default_rng(3).random((16, 3, 256, 192), dtype=np.float32).astype(np.float16)
    a5, a8 = e.hrnet_forward(crops[:5]), e.hrnet_forward(crops[:8])           # capacity 8
    a11 = e.hrnet_forward(crops[:11])                                          # grows the buffers: capacity 16, new plans
    b5, b8, a16 = e.hrnet_forward(crops[:5]), e.hrnet_forward(crops[:8]), e.hrnet_forward(crops)
    assert np.array_equal(a8[:5], a5) and np.array_equal(a5, b5) and np.array_equal(a8, b8)
    assert np.array_equal(a11[:5], b5) and np.array_equal(a16[:11], a11)
    e.close()


def test_hrnet_w48_384x288(eng):
    """reference model size (modules/pose_estimator.py:30, models/conv.py:61)."""
    from oracle.hrnet_fp32 import HRNetFP32
    w = eng.load_hrnet(None, 48, 384, 288, seed=1)
    rng = np.random.default_rng(6)
    crops = rng.uniform(0, 1, (2, 3, 384, 288)).astype(np.float16)
    ref = HRNetFP32(w, 48)(crops.astype(np.float32)).numpy()
    hm = eng.hrnet_forward(crops, np.float32)
    assert hm.shape == (2, 17, 96, 72)
    assert rel_err(hm, ref) < 1e-2
