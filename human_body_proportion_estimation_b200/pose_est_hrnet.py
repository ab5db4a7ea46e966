"""`run_demo_pose_est` (human_body_length_est/pose_est_hrnet_trtserver.py:31-146):
single-person HRNet on whole frames.  The reference returns None and only draws; this returns
[(keypts(17,2) image px, conf(17,1)), ...] per frame.

Documented deviation: the reference's preprocessing (:15-19) divides by 255 FIRST and lets cv2.resize
interpolate the float64 image, and it resizes only when the model has a fixed input size; here the frame is resized as
uint8 with cv2.resize's 11-bit fixed point (HBP_PRE_STRETCH, the arithmetic of PoseEstimator.preprocess, which IS
bit-exact against the reference) and divided by 255 afterwards, always to the model size.  The two differ by the uint8
rounding of the interpolated value, at most 0.5/255 per pixel -- below the fp16 resolution of the HRNet input for half
of the value range and two orders of magnitude below the 1e-2 heatmap tolerance; no golden pins this entry point."""
import numpy as np

from . import engine as _engine
from ._capi import PRE_STRETCH
from .person_det_pose import _load_media


def preprocess(img, width=288, height=384, new_type=np.float32, engine=None):
    """(H,W,3) BGR uint8 -> (3,height,width) RGB in [0,1]"""
    eng = engine or _engine.default_engine()
    return eng.preprocess(np.asarray(img), PRE_STRETCH, height, width, True, 128, new_type)[0]


def run_demo_pose_est(media_filename, model_name="hrnet_w48_384x288", person_height=[175],
                      inference_mode="video", det_threshold=0.55, save_result_dir=None, debug=True,
                      *, frames=None, engine=None):
    import re
    eng = engine or _engine.default_engine()
    m = re.search(r"hrnet_w(32|48)_(\d+)x(\d+)", str(model_name))
    width, ih, iw = (int(m.group(1)), int(m.group(2)), int(m.group(3))) if m else (48, 384, 288)
    if eng.hrnet != (width, ih, iw):
        eng.load_hrnet(None, width, ih, iw)
    if frames is None:
        # _load_media yields RGB (the reference's cv2 BGR after its BGR2RGB swap)
        frames = _load_media(media_filename, inference_mode)
    results = []
    for fr in frames:
        h, w = fr.shape[:2]
        x = eng.preprocess(fr, PRE_STRETCH, ih, iw, False, 128, np.float16)
        hm = eng.hrnet_forward(x, np.float32)
        dec = eng.decode_proportions(hm)
        k = dec["kpts_hm"][0].copy()
        # reference :126-129: scale heatmap coords to the image
        k[:, 0] *= w / hm.shape[3]
        k[:, 1] *= h / hm.shape[2]
        results.append((k, dec["scores"][0].reshape(17, 1)))
    return results
