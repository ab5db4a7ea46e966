"""YOLOv5 detection path (human_body_length_est/obj_det_yolov5_onnx.py:27-36,86-177
and obj_det_yolov5_trtserver.py:30-50) on the B200 engine: letterbox preprocess,
raw-head decode, NMS, scale_coords.  The YOLOv5 backbone itself is an opaque ONNX
artifact in the reference and not part of this build: `detect_onnx`
takes the network as a callable `model(batch_chw_f32) -> outputs`."""
import numpy as np

from . import engine as _engine
from . import onnx_utils
from ._capi import PRE_LETTERBOX, PRE_LETTERBOX_PIL


def preprocess_image(pil_image, in_size=(640, 640), engine=None, resample="bicubic"):
    """obj_det_yolov5_onnx.py:27-36: letterbox (PIL BICUBIC like the reference, pad 128), HWC->CHW
    float32, /255.  resample="bilinear": cv2.resize-exact bilinear instead."""
    eng = engine or _engine.default_engine()
    in_w, in_h = in_size
    mode = PRE_LETTERBOX_PIL if resample == "bicubic" else PRE_LETTERBOX
    return eng.preprocess(np.asarray(pil_image), mode, in_h, in_w, False, 128, np.float32)[0]


def postprocess_decoded(output, conf_thres=0.4, iou_thres=0.5, classes=None, engine=None):
    """obj_det_yolov5_trtserver.py:40-50 -> [det_boxes(n,4) letterbox px, det_scores(n,), det_classes(n,)]"""
    det = onnx_utils.non_max_suppression(np.asarray(output), conf_thres, iou_thres, classes, engine=engine)[0]
    return [det[:, :4], det[:, 4], det[:, 5]]


def detect_onnx(src_path, media_type, threshold=0.6, official=True, onnx_path=None, output_dir=None,
                num_classes=80, *, model=None, frames=None, in_size=(640, 640), engine=None):
    """Same arguments as the reference (:86-93); `threshold` is unused there too
    (hard-coded 0.4/0.5 and 0.4/0.3, :122,172).  Returns the per-frame detection
    lists the reference only draws."""
    if model is None:
        raise RuntimeError("detect_onnx needs `model=` (the YOLOv5 network is not part of this build)")
    eng = engine or _engine.default_engine()
    if frames is None:
        from .person_det_pose import _load_media
        frames = _load_media(src_path, media_type)
    results = []
    for fr in frames:
        x = preprocess_image(fr, in_size, eng)[None]
        outputs = model(x)
        if official and len(outputs) == 4:
            dets = onnx_utils.non_max_suppression(np.asarray(outputs[0]), 0.4, 0.5, engine=eng)
        else:
            heads = list(outputs[1:4]) if len(outputs) == 4 else list(outputs)
            dec = eng.yolo_decode_raw(heads, in_size[0], in_size[1])
            dets = onnx_utils.w_non_max_suppression(dec, num_classes, 0.4, 0.3, engine=eng)
        results.append(dets)
    return results
