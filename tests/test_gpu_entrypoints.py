"""GPU tier (-m gpu): the drop-in Python entry points called the way
the reference's callers call them, checked against the oracle on the engine's own heatmaps.

Reference contracts asserted here (paths relative to human_body_length_est/):
  * run_pdet_pose ........ person_det_pose_edet4_trtserver.py:29-38 (signature), :109-111 ([] on
    no media), :131-134,171,201 (per frame [boxes(n,4) yxyx px, heatmaps(n,17,Hh,Wh), dict_0..]),
    models/conv.py:72-79 + uvicorn_server/server.py:61 (no person: [boxes(0,4), heatmaps(1,..)]),
    :166-168 (person_height[min(i, len-1)]), modules/triton_utils.py:84-87 (raw encoded bytes)
  * run_demo_pose_est .... pose_est_hrnet_trtserver.py:31-37,126-129
  * detect_onnx .......... obj_det_yolov5_onnx.py:86-93,121-122,171-172
  * PoseEstimator ........ modules/pose_estimator.py:20-59,74-99,191-200
"""
import os

import numpy as np
import pytest

from human_body_proportion_estimation_b200 import synth

pytestmark = pytest.mark.gpu

H, W = 540, 960


@pytest.fixture(scope="module")
def eng():
    from human_body_proportion_estimation_b200.engine import Engine
    e = Engine(0)
    e.load_hrnet(None, 48, 384, 288, seed=1)
    yield e
    e.close()


def _detections(n_persons, seed, img_h=H, img_w=W):
    """EfficientDet-style outputs for one frame: persons (class 1.0, score >= .75) mixed with other classes
    and low-score persons that the filter must drop (models/conv.py:22-35)."""
    b, s, c = synth.edet_outputs(1, n_persons, img_h, img_w, seed=seed)
    return b[0], s[0], c[0]


def _expected(det, frame_hw, heatmaps, person_height, det_threshold, max_persons):
    from oracle import detect as od
    from oracle import geometry as og
    h, w = frame_hw
    boxes_n, _ = od.edet_person_filter(det[0], det[1], det[2], det_threshold, h // 17, 0, h, w, max_persons)
    return og.frame_postprocess(boxes_n, heatmaps, h, w, person_height), boxes_n


def _check_entry(entry, det, frame_hw, person_height, det_threshold=0.70, max_persons=3):
    from oracle.geometry import NOT_VISIBLE, SEGMENT_KEYS
    boxes, hm = entry[0], entry[1]
    want, boxes_n = _expected(det, frame_hw, hm, person_height, det_threshold, max_persons)
    n = boxes_n.shape[0]
    assert len(entry) == 2 + n
    assert boxes.dtype == np.float32 and boxes.shape == (n, 4)
    assert np.array_equal(boxes, want[0])                                  # de-normalised yxyx px (:145)
    assert hm.dtype == np.float32 and hm.shape == (n, 17, 96, 72)
    for i in range(n):
        got, exp = entry[2 + i], want[2 + i]
        assert list(got.keys()) == list(SEGMENT_KEYS)
        for k in SEGMENT_KEYS:
            if isinstance(exp[k], str):
                assert got[k] == NOT_VISIBLE, (i, k)
            else:
                # np.float32 segments, np.float64 torso (pose_estimator.py:177-178 on int midpoints)
                assert type(got[k]) is (np.float64 if k == "torso" else np.float32), (i, k, type(got[k]))
                assert float(got[k]) == float(exp[k]), (i, k, got[k], exp[k])


def test_run_pdet_pose_frames_and_height_indexing(eng):
    from human_body_proportion_estimation_b200.person_det_pose import run_pdet_pose
    frames = [synth.frame_u8(H, W, seed=90 + i) for i in range(2)]
    dets = [_detections(3, seed=300 + i) for i in range(2)]
    heights = [181, 166]                      # third person re-uses the last entry (:166-168)
    out = run_pdet_pose(None, person_height=heights, frames=frames, detections=dets, engine=eng, debug=False)
    assert isinstance(out, list) and len(out) == 2
    for entry, det in zip(out, dets):
        assert isinstance(entry, list) and len(entry) == 5
        _check_entry(entry, det, (H, W), heights)
    # reference default max_persons is 3 (models/conv.py:34-35) although more persons pass the filter
    more = _detections(6, seed=310)
    out = run_pdet_pose(None, person_height=[175], frames=frames[:1], detections=[more], engine=eng, debug=False)
    assert len(out[0]) == 2 + 3
    out = run_pdet_pose(None, person_height=[175], frames=frames[:1], detections=[more], engine=eng, debug=False,
                        max_persons=6)
    _check_entry(out[0], more, (H, W), [175], max_persons=6)
    # a detector callable sees the frame the model would see
    seen = []

    def detector(fr):
        seen.append(fr.shape)
        return dets[0]
    out2 = run_pdet_pose(None, frames=frames[:1], detector=detector, engine=eng, debug=False)
    assert seen == [(H, W, 3)]
    _check_entry(out2[0], dets[0], (H, W), [175])


def test_run_pdet_pose_threshold_and_no_person(eng):
    from human_body_proportion_estimation_b200.person_det_pose import run_pdet_pose
    frame = synth.frame_u8(H, W, seed=95)
    det = _detections(2, seed=320)
    # a threshold above every score: the ensemble still answers with one heatmap set of the zero crop
    out = run_pdet_pose(None, frames=[frame], detections=[det], det_threshold=0.999, engine=eng, debug=False)
    assert len(out) == 1 and len(out[0]) == 2
    assert out[0][0].shape == (0, 4) and out[0][0].dtype == np.float32
    assert out[0][1].shape == (1, 17, 96, 72) and out[0][1].dtype == np.float32
    zero = eng.hrnet_forward(np.zeros((1, 3, 384, 288), np.float16), np.float32)
    assert np.array_equal(out[0][1], zero)
    # lower threshold keeps more rows than the default (scores in [0.05, 0.99])
    out = run_pdet_pose(None, frames=[frame], detections=[det], det_threshold=0.30, engine=eng, debug=False)
    _check_entry(out[0], det, (H, W), [175], det_threshold=0.30)


def test_run_pdet_pose_media_forms(eng, tmp_path):
    """path, directory and raw encoded bytes (triton_utils.py:75-128); unreadable media -> []"""
    import cv2
    from human_body_proportion_estimation_b200.person_det_pose import run_pdet_pose
    rgb = synth.frame_u8(H, W, seed=97)
    bgr = np.ascontiguousarray(rgb[..., ::-1])
    p0 = str(tmp_path / "a.png")
    assert cv2.imwrite(p0, bgr)
    det = _detections(2, seed=330)
    via_frames = run_pdet_pose(None, frames=[rgb], detections=[det], engine=eng, debug=False)
    via_path = run_pdet_pose(p0, detections=[det], engine=eng, debug=False)           # cv2 BGR -> RGB (:15-18)
    assert len(via_path) == 1
    assert np.array_equal(via_path[0][1], via_frames[0][1])
    _check_entry(via_path[0], det, (H, W), [175])
    # directory: every file, sorted
    d = tmp_path / "dir"
    os.makedirs(d)
    cv2.imwrite(str(d / "0.png"), bgr)
    cv2.imwrite(str(d / "1.png"), bgr[:, ::-1].copy())
    via_dir = run_pdet_pose(str(d), detections=[det, det], engine=eng, debug=False)
    assert len(via_dir) == 2
    assert np.array_equal(via_dir[0][1], via_frames[0][1])
    assert not np.array_equal(via_dir[1][1], via_frames[0][1])
    # raw bytes: PIL decodes to RGB, the reference's preprocess swaps it once more, so the model sees BGR
    raw = open(p0, "rb").read()
    via_bytes = run_pdet_pose(raw, detections=[det], engine=eng, debug=False)
    swapped = run_pdet_pose(None, frames=[bgr], detections=[det], engine=eng, debug=False)
    assert np.array_equal(via_bytes[0][1], swapped[0][1])
    _check_entry(via_bytes[0], det, (H, W), [175])
    # nothing readable -> [] (reference :109-111)
    assert run_pdet_pose(str(tmp_path / "missing.png"), engine=eng, debug=False) == []
    # model_name / grpc_port are accepted and ignored
    again = run_pdet_pose(p0, "whatever_model", [175], "image", 0.70, None, "1234", False, detections=[det], engine=eng)
    assert np.array_equal(again[0][1], via_path[0][1])


def test_run_pdet_pose_strict_reraises(eng):
    """the reference raises UnboundLocalError when a shoulder / hip is below its gate (pose_estimator.py:146-157);
    the default returns "Part not visible" there, strict=True re-raises"""
    from human_body_proportion_estimation_b200.person_det_pose import run_pdet_pose
    from oracle import geometry as og
    frame = synth.frame_u8(H, W, seed=99)
    det = _detections(3, seed=340)
    out = run_pdet_pose(None, frames=[frame], detections=[det], engine=eng, debug=False)
    # does the oracle's reference-faithful path raise on any of these persons?
    raises = False
    for i in range(len(out[0]) - 2):
        r = og.person_postprocess(out[0][1][i], out[0][0][i], 175)
        if {5, 6, 11, 12} & r["ignored"]:
            raises = True
            d = out[0][2 + i]
            assert d["torso"] == og.NOT_VISIBLE
    if raises:
        with pytest.raises(UnboundLocalError):
            run_pdet_pose(None, frames=[frame], detections=[det], engine=eng, debug=False, strict=True)
    else:          # random-init heatmaps rarely clear every gate; make the case explicit if they did
        out_s = run_pdet_pose(None, frames=[frame], detections=[det], engine=eng, debug=False, strict=True)
        assert len(out_s[0]) == len(out[0])


def test_run_demo_pose_est(eng):
    """single-person HRNet on whole frames: keypoints scaled heatmap -> image (pose_est_hrnet_trtserver.py:126-129)"""
    from human_body_proportion_estimation_b200._capi import PRE_STRETCH
    from human_body_proportion_estimation_b200.pose_est_hrnet import run_demo_pose_est
    from oracle import geometry as og
    frames = [synth.frame_u8(H, W, seed=120 + i) for i in range(2)]
    res = run_demo_pose_est(None, "hrnet_w48_384x288", frames=frames, engine=eng, debug=False)
    assert len(res) == 2
    for fr, (k, conf) in zip(frames, res):
        assert k.shape == (17, 2) and conf.shape == (17, 1)
        x = eng.preprocess(fr, PRE_STRETCH, 384, 288, False, 128, np.float16)
        hm = eng.hrnet_forward(x, np.float32)
        xy, score, _ = og.decode_heatmap(hm[0])
        want = xy.copy()
        want[:, 0] *= W / 72
        want[:, 1] *= H / 96
        assert np.array_equal(k, want)
        assert np.array_equal(conf, score)


def test_detect_onnx_with_stub_model(eng):
    """detect_onnx(model=<callable>): letterbox -> model -> NMS; official (4 outputs, decoded head first) and
    raw-head branches (obj_det_yolov5_onnx.py:107-122,123-172)"""
    from human_body_proportion_estimation_b200.obj_det_yolov5 import detect_onnx, preprocess_image
    from oracle import detect as od
    frame = synth.frame_u8(H, W, seed=130)
    pred, _ = synth.yolo_decoded_head(n_persons=8, n_distract=40, seed=77)
    rng = np.random.default_rng(5)
    heads = [rng.normal(-3.0, 2.0, (1, 3, s, s, 85)).astype(np.float32) for s in (20, 40, 80)]
    seen = []

    def model(x):
        seen.append(x.copy())
        return [pred.copy()] + heads
    out = detect_onnx(None, "image", model=model, frames=[frame], engine=eng)
    assert len(out) == 1 and len(out[0]) == 1
    want = od.official_nms(pred, 0.4, 0.5)[0]
    assert np.array_equal(np.asarray(out[0][0]), want)
    # the network saw exactly the letterboxed tensor preprocess_image builds: (1,3,640,640) f32 in [0,1]
    assert seen[0].shape == (1, 3, 640, 640) and seen[0].dtype == np.float32
    assert np.array_equal(seen[0][0], preprocess_image(frame, (640, 640), eng))
    # raw-head branch: three scales only -> decode + legacy NMS (conf 0.4 / nms 0.3, :171-172)
    out_raw = detect_onnx(None, "image", official=False, model=lambda x: heads, frames=[frame], engine=eng)
    dec = od.yolo_raw_decode(heads, 640, 640, 80)
    want_raw = od.legacy_nms(dec.copy(), 80, 0.4, 0.3)[0]
    got_raw = out_raw[0][0]
    if want_raw is None:
        assert got_raw is None
    else:
        got_raw = np.asarray(got_raw)
        assert got_raw.shape == want_raw.shape
        assert np.allclose(got_raw, want_raw, rtol=1e-5, atol=1e-4)
    with pytest.raises(RuntimeError):
        detect_onnx(None, "image", frames=[frame], engine=eng)


def test_legacy_nms_row_width_contract(eng):
    """rows wider than 5+num_classes are legal in the reference (it slices [:, 5:5+nc]); narrower rows
    must be refused instead of read out of bounds."""
    from oracle import detect as od
    rng = np.random.default_rng(11)
    p = rng.uniform(0, 1, (1, 500, 5 + 12)).astype(np.float32)
    p[..., :2] = rng.uniform(50, 590, (1, 500, 2)); p[..., 2:4] = rng.uniform(10, 120, (1, 500, 2))
    got = eng.yolo_nms_legacy(p.copy(), 8, 0.5, 0.4)[0]
    want = od.legacy_nms(p[..., :13].copy(), 8, 0.5, 0.4)[0]
    assert np.array_equal(got, want)
    with pytest.raises(ValueError):
        eng.yolo_nms_legacy(p.copy(), 80, 0.5, 0.4)


def test_pose_estimator_class(eng):
    """PoseEstimator(model_path).inference / static helpers (modules/pose_estimator.py:20-59,74-99,191-200)"""
    from human_body_proportion_estimation_b200 import engine as E
    from human_body_proportion_estimation_b200.pose_estimator import PoseEstimator
    from oracle import geometry as og
    from oracle import imgproc
    E._default[0] = eng                      # PoseEstimator binds the process-wide engine of its device
    pe = PoseEstimator("hrnet_w32_256x192", device=0, seed=0)
    assert (pe.h, pe.w, pe.c) == (256, 192, 3)
    bgr = [np.ascontiguousarray(synth.frame_u8(300, 200, seed=140 + i)[..., ::-1]) for i in range(2)]
    x = PoseEstimator.preprocess(np.stack(bgr), pe.w, pe.h, eng)
    assert x.shape == (2, 3, 256, 192) and x.dtype == np.float32
    assert np.array_equal(x, imgproc.hrnet_preprocess(np.stack(bgr), pe.w, pe.h))
    hm = pe.inference(np.stack(bgr))
    assert hm.shape == (2, 17, 64, 48) and hm.dtype == np.float32
    assert np.array_equal(hm, eng.hrnet_forward(x.astype(np.float16), np.float32))
    one = pe.inference(bgr[0])               # (H,W,C) form
    assert np.array_equal(one[0], hm[0])
    k, mv = PoseEstimator.get_max_pred_keypts_from_heatmap(hm[0], eng)
    xy, score, _ = og.decode_heatmap(hm[0])
    assert np.array_equal(k, xy) and np.array_equal(mv, score)
    # dist dict on remapped keypoints: same values as the oracle's reference-faithful dict
    box = np.array([20.0, 30.0, 280.0, 170.0], np.float32)
    r = og.person_postprocess(hm[0], box, 170)
    if not ({5, 6, 11, 12} & r["ignored"]):
        d = PoseEstimator.get_keypoint_dist_dict(r["pixel_to_cm"], r["xy_img"], r["ignored"])
        for key, v in r["lengths"].items():
            assert (d[key] == v) if isinstance(v, str) else (float(d[key]) == float(v)), key
    d = PoseEstimator.get_keypoint_dist_dict(0.5, r["xy_img"], {5})
    assert d["shoulder"] == og.NOT_VISIBLE and d["torso"] == og.NOT_VISIBLE
    # hbp_keypoint_lengths on keypoints alone: every value and type equal to the reference-faithful dict
    rng = np.random.default_rng(1)
    for trial in range(8):
        kk = rng.uniform(-50, 800, (17, 2)).astype(np.float32)
        for ign in (set(), {0, 9}, {5}, {11, 13}, {16, 14, 7}):
            a = PoseEstimator.get_keypoint_dist_dict(0.41, kk, ign, engine=eng)
            b = og.lengths_dict(0.41, kk, ign)
            assert a == b and all(type(a[k_]) is type(b[k_]) for k_ in a), (trial, ign)
    batch = eng.keypoint_lengths(np.stack([r["xy_img"]] * 5), np.linspace(0.2, 0.9, 5))
    assert batch["lengths_cm"].shape == (5, 11) and batch["torso_cm"].shape == (5,)
    with pytest.raises(UnboundLocalError):
        PoseEstimator.get_keypoint_dist_dict(0.5, r["xy_img"], {5}, strict=True)
    eng.load_hrnet(None, 48, 384, 288, seed=1)     # restore the module fixture's model
