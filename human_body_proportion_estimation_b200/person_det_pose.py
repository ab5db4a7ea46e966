"""`run_pdet_pose` with the signature and return structure of the reference
(human_body_length_est/person_det_pose_edet4_trtserver.py:29-38, 201), running
in-process on the B200 engine: no Triton server, no gRPC.

Per frame the reference's ensemble did (models/conv.py:14-86): EfficientDet-Lite4
-> person filter + box expansion -> crop_and_resize -> HRNet, then the client
decoded the heatmaps per person (:148-171).  Here: detector outputs ->
hbp_edet_person_filter -> hbp_pose_pipeline (crop -> HRNet -> decode+lengths).

The EfficientDet backbone is not part of this build (the reference ships it only
as a Google-Drive artifact).  Detections therefore come from the
`detector` callable / `detections` argument; without either the whole frame is
taken as one person box so the entry point stays runnable.
"""
import io
import os

import numpy as np

from . import engine as _engine
from . import geometry

_IMG_EXT = (".jpg", ".jpeg", ".png", ".bmp", ".webp")


def _load_media(media_filename, inference_mode):
    """reference modules/triton_utils.py:75-128: path, directory or raw encoded bytes.
    Returns a list of (H,W,3) uint8 arrays exactly as the reference's `preprocess`
    hands them to the model (:15-18: cv2 BGR -> RGB; raw bytes are RGB-decoded and
    then swapped as well, so they reach the model as BGR -- quirk kept)."""
    import cv2
    frames = []
    if isinstance(media_filename, (bytes, bytearray)):
        from PIL import Image
        img = np.asarray(Image.open(io.BytesIO(media_filename)).convert("RGB"))
        return [np.ascontiguousarray(img[..., ::-1])]
    if os.path.isdir(media_filename):
        names = sorted(os.path.join(media_filename, f) for f in os.listdir(media_filename)
                       if os.path.isfile(os.path.join(media_filename, f)))
    else:
        names = [media_filename]
    for nm in names:
        if inference_mode == "video" and not nm.lower().endswith(_IMG_EXT):
            cap = cv2.VideoCapture(nm)
            n = int(cap.get(cv2.CAP_PROP_FRAME_COUNT))
            if n > 10000:       # reference triton_utils.py:100-101
                print("Video must be <10000 frames"); continue
            while True:
                ok, fr = cap.read()
                if not ok:
                    break
                frames.append(np.ascontiguousarray(fr[..., ::-1]))
            cap.release()
        else:
            img = cv2.imread(nm)
            if img is None:
                print(f"failed to load image {nm}"); continue
            frames.append(np.ascontiguousarray(img[..., ::-1]))
    return frames


def run_pdet_pose(media_filename,
                  model_name="ensemble_edet4_person_det_pose",
                  person_height=[175],
                  inference_mode='image',
                  det_threshold=0.70,
                  save_result_dir=None,
                  grpc_port='8994',
                  debug=True,
                  *, detector=None, detections=None, max_persons=3, frames=None,
                  arch="hrnet_w48_384x288", engine=None, strict=False):
    """Returns [[boxes(n,4) yxyx px, heatmaps(n,17,Hh,Wh) f32, dict_0, ...], ...] per frame,
    [] when no media could be read (reference :109-111).  `model_name`/`grpc_port`
    are accepted and ignored.  Extra keyword-only arguments:
      detector(frame_rgb_u8) -> (boxes(K,4) yxyx px, scores(K), classes(K)), or
      detections = [that tuple per frame]; frames = preloaded list of RGB uint8 frames."""
    eng = engine or _engine.default_engine()
    if frames is None:
        frames = _load_media(media_filename, inference_mode)
    if len(frames) == 0:
        print("Image data was missing")
        return []
    import re
    m = re.search(r"hrnet_w(32|48)_(\d+)x(\d+)", arch)
    width, ih, iw = int(m.group(1)), int(m.group(2)), int(m.group(3))
    if eng.hrnet != (width, ih, iw):
        eng.load_hrnet(None, width, ih, iw)
    out = []
    for fi, frame in enumerate(frames):
        h, w = frame.shape[:2]
        if detections is not None:
            det = detections[fi]
        elif detector is not None:
            det = detector(frame)
        else:
            det = (np.array([[0, 0, h, w]], np.float32), np.array([1.0], np.float32), np.array([1.0], np.float32))
        # reference :116-117: x_expand = image_height // 17 (sic), y_expand = 0
        x_expand, y_expand = h // 17, 0
        boxes_n = eng.edet_person_filter(det[0], det[1], det[2], det_threshold, x_expand, y_expand, h, w,
                                         max_persons=max_persons)[0]
        n = boxes_n.shape[0]
        if n == 0:
            # models/conv.py:72-79: zero crop -> the ensemble still returns one heatmap set
            hm = eng.hrnet_forward(np.zeros((1, 3, ih, iw), np.float16), np.float32)
            out.append([np.zeros((0, 4), np.float32), hm])
            continue
        boxes_px = boxes_n.copy()
        boxes_px *= [h, w, h, w]                                       # reference :145
        mats = geometry.crop_and_resize_matrices(boxes_n, h, w, ih, iw)
        hcm = [person_height[min(i, len(person_height) - 1)] for i in range(n)]
        # the frame already is what the reference feeds the model; no further swap
        res = eng.pose_pipeline(frame, mats, np.zeros(n, np.int32), boxes_px, hcm, swap_rb=False,
                                return_heatmaps=np.float32)
        entry = [boxes_px, res["heatmaps"]]
        for i in range(n):
            if strict and (int(res["ignored"][i]) & 0b1100001100000):
                raise UnboundLocalError("chest/crotch unbound (reference pose_estimator.py:146-157)")
            entry.append(_engine.lengths_to_dict(res["lengths_cm"][i], res["torso_cm"][i]))
        out.append(entry)
        if debug:
            print(f"frame {fi}: {n} person(s)")
    return out
