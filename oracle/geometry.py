"""Oracle (TEST INFRASTRUCTURE): heatmap decode, box remap, score gate and
body-part lengths, restated in numpy from the reference.

Reference anchors (paths relative to /root/reference/human_body_length_est):
  * decode ............ modules/pose_estimator.py:74-99
  * remap + gate ...... person_det_pose_edet4_trtserver.py:145-168
  * segments/lengths .. modules/pose_estimator.py:130-200

All arithmetic is written so that every intermediate has the dtype numpy 2.x
gives the reference code (float32 keypoints, float64 torso, float32 scaling by
a weak python float) -- the fixtures in tests/golden/decode_*.npz were produced
by the reference itself and this file is checked against them bit for bit.
"""
import numpy as np

# modules/pose_estimator.py:9-17 -- COCO-17 order, viewer's left/right.
JOINT_NAMES = ("nose", "reye", "leye", "rear", "lear", "rshoulder", "lshoulder",
               "relbow", "lelbow", "rwrist", "lwrist", "rhip", "lhip",
               "rknee", "lknee", "rankle", "lankle")

# person_det_pose_edet4_trtserver.py:62-63 -- per-joint score gates.
JOINT_THRESHOLDS = (0.45, 0.46, 0.45, 0.40, 0.34, 0.10, 0.10, 0.10, 0.10,
                    0.24, 0.30, 0.11, 0.10, 0.15, 0.10, 0.25, 0.20)

# modules/pose_estimator.py:156-166 -- output key -> (joint a, joint b); the
# keys are from the person's point of view, the joints from the viewer's,
# hence the mirror.  -1 = chest (int midpoint of 5,6), -2 = crotch (11,12).
CHEST, CROTCH = -1, -2
SEGMENTS = (
    ("shoulder", 5, 6),
    ("torso", CROTCH, CHEST),
    ("lshoulder_lelbow", 5, 7),
    ("rshoulder_relbow", 6, 8),
    ("lwrist_lelbow", 9, 7),
    ("rwrist_relbow", 10, 8),
    ("rhip_lhip", 12, 11),
    ("rhip_rknee", 12, 14),
    ("lhip_lknee", 11, 13),
    ("rankle_rknee", 16, 14),
    ("lankle_lknee", 15, 13),
)
SEGMENT_KEYS = tuple(s[0] for s in SEGMENTS)
NOT_VISIBLE = "Part not visible"


def decode_heatmap(hm):
    """pose_estimator.py:74-99.  hm (J,H,W) -> keypts (J,2) f32 (x,y), score (J,1).

    First-index argmax, coordinates zeroed unless max > 0 (NaN -> (0,0) with a
    NaN score, all-negative -> (0,0) with the negative max as score).
    Also returns the flat argmax index (int64) for the bit-exact index check.
    """
    hm = np.asarray(hm)
    J, H, W = hm.shape
    flat = hm.reshape(J, H * W)
    idx = np.argmax(flat, axis=1)
    score = np.max(flat, axis=1).reshape(J, 1)
    fidx = idx.astype(np.float32)
    xy = np.empty((J, 2), np.float32)
    xy[:, 0] = fidx % np.float32(W)
    xy[:, 1] = np.floor(fidx / np.float32(W))
    xy *= np.greater(score, 0.0).astype(np.float32)
    return xy, score, idx


def quarter_offset(hm, xy):
    """Optional sub-pixel step of the public HRNet `get_final_preds` (NOT in the
    reference -- SURVEY.md A.2): +-0.25 px toward the larger neighbour, only for
    strictly interior maxima."""
    J, H, W = hm.shape
    out = xy.copy()
    for j in range(J):
        x, y = int(xy[j, 0]), int(xy[j, 1])
        if 1 < x < W - 1 and 1 < y < H - 1:
            dx = hm[j, y, x + 1] - hm[j, y, x - 1]
            dy = hm[j, y + 1, x] - hm[j, y - 1, x]
            out[j, 0] += np.float32(0.25) * np.sign(dx).astype(np.float32)
            out[j, 1] += np.float32(0.25) * np.sign(dy).astype(np.float32)
    return out


def remap_to_image(xy, box_yxyx_px, hm_h, hm_w):
    """person_det_pose_edet4_trtserver.py:151-160.

    The box (already multiplied by [h,w,h,w]) is truncated with int(); the
    keypoints are divided by the heatmap size, multiplied by the truncated crop
    size and shifted by the truncated corner.  numpy evaluates the three in-place
    steps in float64 and rounds to float32 after each (int64 list operands); for
    one IEEE operation that equals the float32 result, so float32 ops are used.
    Returns (xy_img f32 (J,2), (x1,y1,x2,y2) ints).
    """
    x1, y1 = int(box_yxyx_px[1]), int(box_yxyx_px[0])
    x2, y2 = int(box_yxyx_px[3]), int(box_yxyx_px[2])
    k = np.array(xy, np.float32, copy=True)
    k /= [hm_w, hm_h]
    k *= [x2 - x1, y2 - y1]
    k += [x1, y1]
    return k, (x1, y1, x2, y2)


def ignored_joints(score, thresholds=JOINT_THRESHOLDS):
    """person_det_pose_edet4_trtserver.py:162-163: joint j ignored iff
    score_j < T_j, compared in float32 (weak python float)."""
    s = np.asarray(score, np.float32).reshape(-1)
    t = np.asarray(thresholds, np.float32)
    return {j for j in range(s.shape[0]) if s[j] < t[j]}


def _int_mid(a, b):
    # pose_estimator.py:147-152: int(a + b) // 2 on the float32 sum.
    return int(np.float32(a) + np.float32(b)) // 2


def segment_lengths_px(xy_img, ignored, strict=False):
    """pose_estimator.py:130-180.  Returns a list of 11 values in SEGMENTS
    order: np.float32 norm (np.float64 for the torso, whose endpoints are python
    ints) or 0 when an endpoint is ignored.

    The reference raises UnboundLocalError when a shoulder or hip is ignored
    (chest / crotch only bound conditionally, :146-157).  strict=True re-raises
    the same way; the default treats an unbound chest/crotch as "missing".
    """
    k = np.asarray(xy_img, np.float32)
    ign = set(ignored) if ignored is not None else set()
    have_chest = 5 not in ign and 6 not in ign
    have_crotch = 11 not in ign and 12 not in ign
    if strict and not (have_chest and have_crotch):
        raise UnboundLocalError("chest/crotch referenced before assignment "
                                "(reference pose_estimator.py:146-157)")
    chest = [_int_mid(k[5, 0], k[6, 0]), _int_mid(k[5, 1], k[6, 1])] if have_chest else None
    crotch = [_int_mid(k[11, 0], k[12, 0]), _int_mid(k[11, 1], k[12, 1])] if have_crotch else None

    def point(j):
        if j == CHEST:
            return chest
        if j == CROTCH:
            return crotch
        return None if j in ign else k[j]

    out = []
    for _, a, b in SEGMENTS:
        pa, pb = point(a), point(b)
        if pa is None or pb is None:
            out.append(0)
        else:
            out.append(np.linalg.norm(np.asarray(pa) - np.asarray(pb)))
    return out


def lengths_dict(pixel_to_cm, xy_img, ignored, strict=False):
    """pose_estimator.py:191-200: value*pixel_to_cm when value > 0 else the
    string "Part not visible"."""
    vals = segment_lengths_px(xy_img, ignored, strict=strict)
    return {key: (v * pixel_to_cm if v > 0 else NOT_VISIBLE)
            for key, v in zip(SEGMENT_KEYS, vals)}


def person_postprocess(hm, box_yxyx_px, height_cm, thresholds=JOINT_THRESHOLDS,
                       strict=False, quarter=False):
    """One iteration of the per-person loop of run_pdet_pose
    (person_det_pose_edet4_trtserver.py:148-171)."""
    J, H, W = hm.shape
    xy, score, idx = decode_heatmap(hm)
    if quarter:
        xy = quarter_offset(hm, xy)
    xy_img, (x1, y1, x2, y2) = remap_to_image(xy, box_yxyx_px, H, W)
    ign = ignored_joints(score, thresholds)
    pixel_to_cm = height_cm / (y2 - y1)
    d = lengths_dict(pixel_to_cm, xy_img, ign, strict=strict)
    return dict(xy_hm=xy, score=score, idx=idx, xy_img=xy_img, ignored=ign,
                pixel_to_cm=pixel_to_cm, lengths=d)


def frame_postprocess(boxes_norm_yxyx, heatmaps, img_h, img_w, person_height=(175,),
                      thresholds=JOINT_THRESHOLDS, strict=False):
    """run_pdet_pose's per-response body (:133-171): returns
    [boxes_px, heatmaps, dict_0, ...] like box_hmap_list[-1]."""
    boxes = np.array(boxes_norm_yxyx, copy=True)
    boxes *= [img_h, img_w, img_h, img_w]
    out = [boxes, heatmaps]
    for i, (hm, box) in enumerate(zip(heatmaps, boxes)):
        h_cm = person_height[min(i, len(person_height) - 1)]
        out.append(person_postprocess(hm, box, h_cm, thresholds, strict)["lengths"])
    return out


def lengths_to_array(d):
    """dict -> (11,) float64 with 0.0 for "Part not visible" (the C-ABI's
    convention for hbp_decode_proportions)."""
    return np.array([0.0 if isinstance(d[k], str) else float(d[k]) for k in SEGMENT_KEYS])
