"""Detector post-processing with the reference's function names
(human_body_length_est/modules/onnx_utils.py), on the B200 engine.

Inputs/outputs are torch tensors when torch tensors come in (the reference's
types), numpy otherwise.
"""
import numpy as np

from . import engine as _engine
from ._capi import PRE_LETTERBOX, PRE_LETTERBOX_PIL, NHWC


def _np(x):
    return x.detach().cpu().numpy() if hasattr(x, "detach") else np.asarray(x)


def _like(ref, arr):
    if hasattr(ref, "detach"):
        import torch
        return torch.from_numpy(arr)
    return arr


def non_max_suppression(prediction, conf_thres=0.25, iou_thres=0.45, classes=None, agnostic=False,
                        multi_label=False, labels=(), engine=None):
    """onnx_utils.py:125-222 -> list of (n,6) [xyxy, conf, cls] per image.  Only the
    configuration the reference's callers use is built (best class, class-offset NMS)."""
    assert 0 <= conf_thres <= 1, f'Invalid Confidence threshold {conf_thres}, valid values are between 0.0 and 1.0'
    assert 0 <= iou_thres <= 1, f'Invalid IoU {iou_thres}, valid values are between 0.0 and 1.0'
    if agnostic or multi_label or (labels and len(labels)):
        raise NotImplementedError("agnostic / multi_label / labels are unused by every reference caller")
    eng = engine or _engine.default_engine()
    out = eng.yolo_nms(_np(prediction), conf_thres, iou_thres, classes)
    return [_like(prediction, o) for o in out]


def w_non_max_suppression(prediction, num_classes, conf_thres=0.5, nms_thres=0.4, engine=None):
    """onnx_utils.py:39-95 -> list of (n,7) or None.  Like the reference, rewrites
    prediction[..., :4] from xywh to corners IN PLACE (:42-47)."""
    eng = engine or _engine.default_engine()
    p = _np(prediction)
    out = eng.yolo_nms_legacy(p, num_classes, conf_thres, nms_thres)
    half_w, half_h = p[..., 2] / 2, p[..., 3] / 2
    corners = np.stack([p[..., 0] - half_w, p[..., 1] - half_h, p[..., 0] + half_w, p[..., 1] + half_h], -1)
    prediction[..., :4] = _like(prediction, corners.astype(np.float32))
    return [None if o is None else _like(prediction, o) for o in out]


def letterbox_image(image, size, engine=None, resample="bicubic"):
    """onnx_utils.py:225-235: aspect-keeping resize with PIL's antialiased BICUBIC (reproduced bit
    for bit on the GPU), centred paste on grey 128.  Takes/returns a PIL image or an (H,W,3) uint8
    array.  resample="bilinear" selects the cv2.resize-exact bilinear sampler instead."""
    eng = engine or _engine.default_engine()
    arr = np.asarray(image)
    w, h = size
    mode = PRE_LETTERBOX_PIL if resample == "bicubic" else PRE_LETTERBOX
    out = eng.preprocess(arr, mode, h, w, False, 128, np.uint8, NHWC)[0]
    if not isinstance(image, np.ndarray):
        from PIL import Image
        return Image.fromarray(out)
    return out


def clip_coords(boxes, img_shape):
    """onnx_utils.py:238-249 (in place)"""
    b = boxes
    if hasattr(b, "clamp_"):
        b[:, 0].clamp_(0, img_shape[1]); b[:, 1].clamp_(0, img_shape[0])
        b[:, 2].clamp_(0, img_shape[1]); b[:, 3].clamp_(0, img_shape[0])
    else:
        b[:, [0, 2]] = np.clip(b[:, [0, 2]], 0, img_shape[1])
        b[:, [1, 3]] = np.clip(b[:, [1, 3]], 0, img_shape[0])


def scale_coords(img1_shape, coords, img0_shape, ratio_pad=None, engine=None):
    """onnx_utils.py:252-266: letterbox px -> original px, in place, returns coords."""
    if ratio_pad is not None:
        raise NotImplementedError("ratio_pad is unused by every reference caller")
    eng = engine or _engine.default_engine()
    if hasattr(coords, "detach"):
        arr = coords.detach().cpu().numpy().astype(np.float32)
        eng.scale_coords(img1_shape, arr, img0_shape)
        coords[:, :4] = _like(coords, arr[:, :4])
        return coords
    return eng.scale_coords(img1_shape, coords, img0_shape)


def xyxy2xywh(x):
    y = x.clone() if hasattr(x, "clone") else np.copy(x)
    y[:, 0] = (x[:, 0] + x[:, 2]) / 2
    y[:, 1] = (x[:, 1] + x[:, 3]) / 2
    y[:, 2] = x[:, 2] - x[:, 0]
    y[:, 3] = x[:, 3] - x[:, 1]
    return y


def xywh2xyxy(x):
    y = x.clone() if hasattr(x, "clone") else np.copy(x)
    y[:, 0] = x[:, 0] - x[:, 2] / 2
    y[:, 1] = x[:, 1] - x[:, 3] / 2
    y[:, 2] = x[:, 0] + x[:, 2] / 2
    y[:, 3] = x[:, 1] + x[:, 3] / 2
    return y
