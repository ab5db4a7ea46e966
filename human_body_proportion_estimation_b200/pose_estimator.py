"""`PoseEstimator` with the reference's method names
(human_body_length_est/modules/pose_estimator.py:20-200), running on the B200
engine instead of onnxruntime + numpy.

Differences from the reference, all deliberate:
  * `model_path` is an architecture name ("hrnet_w32_256x192", "hrnet_w48_384x288",
    ...) or a dict of folded weights -- the ONNX artifacts are not available.
  * `_get_kp_dict` / `get_keypoint_dist_dict` do not raise UnboundLocalError
    when a shoulder or hip is ignored (reference :146-157); the segments that
    need the missing chest/crotch come back as "Part not visible".  Pass
    strict=True to get the reference's exception.
  * drawing / plotting helpers are out of scope.
"""
import re

import numpy as np

from . import engine as _engine
from ._capi import PRE_STRETCH

IDX_TO_KEYPOINTS = {
    0: "nose", 1: "reye", 2: "leye", 3: "rear", 4: "lear", 5: "rshoulder", 6: "lshoulder",
    7: "relbow", 8: "lelbow", 9: "rwrist", 10: "lwrist", 11: "rhip", 12: "lhip",
    13: "rknee", 14: "lknee", 15: "rankle", 16: "lankle"}

_ARCH = re.compile(r"hrnet_w(32|48)_(\d+)x(\d+)")


class PoseEstimator:
    def __init__(self, model_path="hrnet_w32_256x192", device=0, seed=0):
        self.engine = _engine.default_engine(device)
        weights = None
        if isinstance(model_path, dict):
            weights, name = model_path, model_path.get("__arch__", "hrnet_w32_256x192")
        else:
            name = str(model_path)
        m = _ARCH.search(name)
        if not m:
            raise ValueError("model_path must name an architecture like hrnet_w32_256x192, got %r" % (model_path,))
        width, self.h, self.w = int(m.group(1)), int(m.group(2)), int(m.group(3))
        self.b, self.c = -1, 3
        self.input_name = "input"
        if weights is not None:
            weights = {k: v for k, v in weights.items() if k != "__arch__"}
        self.weights = self.engine.load_hrnet(weights, width, self.h, self.w, seed)

    @staticmethod
    def preprocess(frame_s, w=288, h=384, engine=None) -> np.ndarray:
        """(B,H,W,C) / (H,W,C) / list of BGR uint8 frames -> (B,3,h,w) float32 RGB in [0,1]
        (reference :29-45: BGR2RGB, cv2.resize, /255.0, CHW)."""
        eng = engine or _engine.default_engine()
        if isinstance(frame_s, list):
            return np.concatenate([eng.preprocess(np.asarray(f), PRE_STRETCH, h, w, True, 128, np.float32)
                                   for f in frame_s])
        frame_s = np.asarray(frame_s)
        if frame_s.ndim == 3:
            frame_s = frame_s[None]
        return eng.preprocess(frame_s, PRE_STRETCH, h, w, True, 128, np.float32)

    def inference(self, frame_s) -> np.ndarray:
        """(B,H,W,C) or (H,W,C) BGR uint8 -> heatmaps (B,17,h/4,w/4) float32 (reference :47-59)"""
        x = PoseEstimator.preprocess(frame_s, self.w, self.h, self.engine)
        return self.engine.hrnet_forward(x.astype(np.float16), np.float32)

    @staticmethod
    def get_max_pred_keypts_from_heatmap(heatmap, engine=None) -> tuple:
        """(J,H,W) -> keypts (J,2) float32 (x,y), maxvals (J,1)  (reference :74-99)"""
        eng = engine or _engine.default_engine()
        hm = np.asarray(heatmap)
        out = eng.decode_proportions(hm[None])
        return out["kpts_hm"][0], out["scores"][0].reshape(-1, 1).astype(hm.dtype if hm.dtype == np.float32 else np.float32)

    @staticmethod
    def get_keypoint_dist_dict(pixel_to_cm, keypts, ignored_kp_idx=None, strict=False, engine=None):
        """Reference :191-200 on ALREADY REMAPPED keypoints: the segment arithmetic of the decode kernel
        (hbp_keypoint_lengths) on keypoints the caller holds.  strict=True reproduces the reference's UnboundLocalError
        when a chest/crotch joint is ignored (reference :146-157)."""
        ign = set(int(j) for j in ignored_kp_idx) if ignored_kp_idx is not None else set()
        if strict and ({5, 6, 11, 12} & ign):
            raise UnboundLocalError("cannot access local variable 'chest'/'crotch' (reference pose_estimator.py:146-157)")
        eng = engine or _engine.default_engine()
        mask = np.uint32(sum(1 << j for j in ign if 0 <= j < 17))
        res = eng.keypoint_lengths(np.asarray(keypts, np.float32)[None], float(pixel_to_cm), np.array([mask], np.uint32))
        out = {}
        for i, key in enumerate(_engine.SEGMENT_KEYS):
            # chest-crotch is python float in the reference (integer midpoints -> float64 norm)
            v = res["torso_cm"][0] if i == 1 else res["lengths_cm"][0, i]
            out[key] = v if v > 0 else _engine.NOT_VISIBLE
        return out
