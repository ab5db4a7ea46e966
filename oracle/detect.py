"""Oracle (TEST INFRASTRUCTURE): detector-head post-processing restated in numpy.

Reference anchors (relative to /root/reference):
  * YOLOv5 raw-head decode ... human_body_length_est/obj_det_yolov5_onnx.py:123-169
  * official NMS ............. human_body_length_est/modules/onnx_utils.py:125-222
      -> torchvision.ops.nms (third party, unpinned; CPU kernel behaviour is
         restated in `greedy_nms` and pinned by tests/golden/nms_*.npz, which
         were produced by the reference calling torchvision 0.26 here)
  * legacy NMS ............... human_body_length_est/modules/onnx_utils.py:8-95
  * scale/clip coords ........ human_body_length_est/modules/onnx_utils.py:238-266
  * EfficientDet person filter models/conv.py:22-57  (TensorFlow graph; parity
      unpinned -- tensorflow is absent, restated from the call site)

All float work is float32 with one rounding per operation, in the reference's
operation order.
"""
import numpy as np

F = np.float32

# obj_det_yolov5_onnx.py:130-131 -- indexed by OUTPUT order (stride 32, 16, 8).
YOLO_ANCHORS = ((116, 90, 156, 198, 373, 326),
                (30, 61, 62, 45, 59, 119),
                (10, 13, 16, 30, 33, 23))


def _sigmoid32(x):
    x = np.asarray(x, F)
    return (F(1) / (F(1) + np.exp(-x, dtype=F))).astype(F)


def yolo_raw_decode(heads, in_w=640, in_h=640, num_classes=80):
    """obj_det_yolov5_onnx.py:133-169.  heads: 3 arrays (B,3,S,S,5+nc) f32 in
    output order -> (B, sum 3*S*S, 5+nc) f32 rows [cx,cy,w,h,obj,cls...].
    The reference reads feature_w from shape[2] and feature_h from shape[3]
    (swapped names, square maps in practice)."""
    outs = []
    for level, out in enumerate(heads):
        out = np.asarray(out, F)
        B = out.shape[0]
        fw, fh = out.shape[2], out.shape[3]
        sw, sh = int(in_w / fw), int(in_h / fh)
        gx, gy = np.meshgrid(np.arange(fw), np.arange(fh))
        anc = np.asarray(YOLO_ANCHORS[level], F).reshape(1, 3, 1, 1, 2)
        s = _sigmoid32(out[..., :4])
        box = np.empty(out[..., :4].shape, F)
        # (sigmoid*2 - 0.5 + grid) * stride; grid is int64 -> torch promotes to f32
        box[..., 0] = ((s[..., 0] * F(2.0) - F(0.5)) + gx.astype(F)) * F(sw)
        box[..., 1] = ((s[..., 1] * F(2.0) - F(0.5)) + gy.astype(F)) * F(sh)
        t = s[..., 2:4] * F(2)
        box[..., 2:4] = (t * t) * anc
        conf = _sigmoid32(out[..., 4])
        cls = _sigmoid32(out[..., 5:])
        outs.append(np.concatenate([box.reshape(B, -1, 4), conf.reshape(B, -1, 1),
                                    cls.reshape(B, -1, num_classes)], -1))
    return np.concatenate(outs, 1)


def xywh_to_xyxy(b):
    """onnx_utils.py:280-288."""
    b = np.asarray(b, F)
    o = np.empty_like(b)
    hw, hh = b[:, 2] / F(2), b[:, 3] / F(2)
    o[:, 0] = b[:, 0] - hw
    o[:, 1] = b[:, 1] - hh
    o[:, 2] = b[:, 0] + hw
    o[:, 3] = b[:, 1] + hh
    return o


def greedy_nms(boxes, scores, iou_thres):
    """torchvision.ops.nms CPU kernel semantics: stable sort by score
    descending, areas (x2-x1)*(y2-y1), suppress j when
    inter/(area_i+area_j-inter) > thr (float32 ratio promoted to double for the
    compare, so NaN never suppresses).  Returns kept indices, score order."""
    boxes = np.asarray(boxes, F)
    scores = np.asarray(scores, F)
    n = boxes.shape[0]
    order = np.argsort(-scores.astype(np.float64), kind="stable")
    x1, y1, x2, y2 = (boxes[:, i] for i in range(4))
    area = (x2 - x1) * (y2 - y1)
    dead = np.zeros(n, bool)
    keep = []
    thr = float(iou_thres)
    with np.errstate(invalid="ignore", divide="ignore"):
        for a in range(n):
            i = order[a]
            if dead[i]:
                continue
            keep.append(i)
            rest = order[a + 1:]
            w = np.maximum(F(0), np.minimum(x2[i], x2[rest]) - np.maximum(x1[i], x1[rest]))
            h = np.maximum(F(0), np.minimum(y2[i], y2[rest]) - np.maximum(y1[i], y1[rest]))
            inter = (w * h).astype(F)
            ovr = inter / ((area[i] + area[rest]) - inter)
            dead[rest[ovr.astype(np.float64) > thr]] = True
    return np.asarray(keep, np.int64)


def yolo_candidates(pred, conf_thres, classes=None):
    """onnx_utils.py:133,155,171-187 for ONE image.  pred (N,5+nc) f32 ->
    (rows (n,6) f32 [x1,y1,x2,y2,conf,cls], source row index (n,))."""
    pred = np.asarray(pred, F)
    ct = F(conf_thres)
    src = np.nonzero(pred[:, 4] > ct)[0]
    x = pred[src]
    if x.shape[0] == 0:
        return np.zeros((0, 6), F), src
    cls = (x[:, 5:] * x[:, 4:5]).astype(F)
    box = xywh_to_xyxy(x[:, :4])
    j = np.argmax(cls, 1)          # first maximal class, like torch.max(1)
    conf = cls[np.arange(cls.shape[0]), j]
    rows = np.concatenate([box, conf[:, None], j[:, None].astype(F)], 1)
    m = conf > ct
    rows, src = rows[m], src[m]
    if classes is not None:
        m = np.isin(rows[:, 5], np.asarray(classes, F))
        rows, src = rows[m], src[m]
    return rows, src


def official_nms(prediction, conf_thres=0.25, iou_thres=0.45, classes=None,
                 max_det=300, max_nms=30000, max_wh=4096):
    """onnx_utils.py:125-222 (best-class, non-agnostic, no merge -- the only
    configuration any reference caller uses).  prediction (B,N,5+nc) f32 ->
    list of (n,6) f32."""
    out = []
    for pred in np.asarray(prediction, F):
        rows, _ = yolo_candidates(pred, conf_thres, classes)
        if rows.shape[0] == 0:
            out.append(np.zeros((0, 6), F))
            continue
        if rows.shape[0] > max_nms:
            rows = rows[np.argsort(-rows[:, 4].astype(np.float64), kind="stable")[:max_nms]]
        off = rows[:, 5:6] * F(max_wh)
        keep = greedy_nms(rows[:, :4] + off, rows[:, 4], iou_thres)[:max_det]
        out.append(rows[keep])
    return out


def legacy_iou_plus1(a, b):
    """onnx_utils.py:8-36 with x1y1x2y2=True: the +1 pixel convention and the
    +1e-16 in the denominator (a no-op in float32 unless the sum is tiny)."""
    ix1, iy1 = np.maximum(a[0], b[:, 0]), np.maximum(a[1], b[:, 1])
    ix2, iy2 = np.minimum(a[2], b[:, 2]), np.minimum(a[3], b[:, 3])
    inter = np.maximum((ix2 - ix1) + F(1), F(0)) * np.maximum((iy2 - iy1) + F(1), F(0))
    aa = ((a[2] - a[0]) + F(1)) * ((a[3] - a[1]) + F(1))
    ab = ((b[:, 2] - b[:, 0]) + F(1)) * ((b[:, 3] - b[:, 1]) + F(1))
    return (inter / (((aa + ab) - inter) + F(1e-16))).astype(F)


def legacy_nms(prediction, num_classes, conf_thres=0.5, nms_thres=0.4):
    """onnx_utils.py:39-95.  prediction (B,N,5+nc) f32 rows [cx,cy,w,h,obj,cls..].
    Returns list of (n,7) f32 [x1,y1,x2,y2,obj,cls_conf,cls] or None per image.
    (The reference additionally overwrites prediction[..., :4] with the corner
    form in place, :47 -- `mutate=True` callers emulate that themselves.)"""
    pred = np.array(prediction, F, copy=True)
    pred[..., :4] = np.stack([xywh_to_xyxy(p[:, :4]) for p in pred])
    out = [None] * pred.shape[0]
    for bi, p in enumerate(pred):
        p = p[p[:, 4] >= F(conf_thres)]
        if p.shape[0] == 0:
            continue
        cc = p[:, 5:5 + num_classes]
        cj = np.argmax(cc, 1)
        det = np.concatenate([p[:, :5], cc[np.arange(len(cj)), cj][:, None],
                              cj[:, None].astype(F)], 1)
        for c in np.unique(det[:, -1]):
            d = det[det[:, -1] == c]
            d = d[np.argsort(-d[:, 4].astype(np.float64), kind="stable")]
            kept = []
            while d.shape[0]:
                kept.append(d[0])
                if d.shape[0] == 1:
                    break
                iou = legacy_iou_plus1(d[0], d[1:])
                d = d[1:][iou < F(nms_thres)]
            kept = np.stack(kept)
            out[bi] = kept if out[bi] is None else np.concatenate([out[bi], kept])
    return out


def scale_coords(img1_shape, coords, img0_shape):
    """onnx_utils.py:252-266 (+ clip_coords :238-249).  coords (n,4) xyxy in the
    letterbox frame img1 (h,w) -> original frame img0 (h,w); python-float gain
    and pad applied to float32 coords (so each step rounds to float32).
    Returns a new array (the reference mutates in place)."""
    c = np.array(coords, F, copy=True)
    gain = max(img1_shape) / max(img0_shape)
    pad_x = (img1_shape[1] - img0_shape[1] * gain) / 2
    pad_y = (img1_shape[0] - img0_shape[0] * gain) / 2
    c[:, [0, 2]] -= F(pad_x)
    c[:, [1, 3]] -= F(pad_y)
    c[:, :4] /= F(gain)
    c[:, [0, 2]] = np.clip(c[:, [0, 2]], 0, img0_shape[1])
    c[:, [1, 3]] = np.clip(c[:, [1, 3]], 0, img0_shape[0])
    return c


def letterbox_geometry(iw, ih, w, h):
    """onnx_utils.py:225-235: scale, resized size and paste offset."""
    scale = min(w / iw, h / ih)
    nw, nh = int(iw * scale), int(ih * scale)
    return scale, nw, nh, (w - nw) // 2, (h - nh) // 2


def edet_person_filter(det_boxes, det_scores, det_classes, det_thres, x_expand,
                       y_expand, img_h, img_w, max_persons=3, person_class=1.0):
    """models/conv.py:22-57.  det_boxes (100,4) yxyx px, scores (100,), classes
    (100,) -> filtered boxes (n,4) yxyx NORMALISED f32, n <= max_persons, in
    detector order.  (reference: max_persons fixed at 3, :34-35)"""
    b = np.asarray(det_boxes, F)
    s = np.asarray(det_scores, F)
    c = np.asarray(det_classes, F)
    sel = np.nonzero(c == F(person_class))[0]
    sel = sel[s[sel] >= F(det_thres)][:max_persons]
    b = b[sel]
    hf, wf = F(img_h), F(img_w)
    y1 = np.clip(b[:, 0] - F(y_expand), F(0), hf)
    x1 = np.clip(b[:, 1] - F(x_expand), F(0), wf)
    y2 = np.clip(b[:, 2] + F(y_expand), F(0), hf)
    x2 = np.clip(b[:, 3] + F(x_expand), F(0), wf)
    return (np.stack([y1, x1, y2, x2], 1) / np.array([hf, wf, hf, wf], F)).astype(F), sel
