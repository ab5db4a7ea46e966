#!/usr/bin/env python
"""Small end-to-end run for compute-sanitizer: the chained det -> pose pipeline (YOLO head ->
NMS -> crop -> HRNet-W32 -> decode) on a 540x960 frame with 8 person slots, then the EfficientDet chain on two frames,
then one frame through hbp_pose_pipeline.  Prints 'sanitize run ok'."""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from human_body_proportion_estimation_b200 import geometry, synth
from human_body_proportion_estimation_b200.engine import Engine

eng = Engine(0)
eng.load_hrnet(None, 32, 256, 192, seed=0)
H, W = 540, 960
frame = synth.frame_u8(H, W, seed=5)
pred, _ = synth.yolo_decoded_head(n_persons=6, n_distract=20, N=2100, seed=9)
for _ in range(2):
    out = eng.det_pose_collect(eng.det_pose_submit_yolo(frame, pred, persons_cap=8, cand_cap=512))
print("yolo chain persons", out["n"], "status", out["status"])
frames = np.stack([frame, synth.frame_u8(H, W, seed=6)])
b, s, c = synth.edet_outputs(2, 3, H, W)
out = eng.det_pose_collect(eng.det_pose_submit_edet(frames, b, s, c, max_persons=4))
print("edet chain persons", out["n"], "status", out["status"])
boxes = synth.person_boxes_yxyx_px(3, H, W, seed=4, hmin=200, hmax=500)
mats = geometry.crop_and_resize_matrices(boxes / np.array([H, W, H, W], np.float32), H, W, 256, 192)
res = eng.pose_pipeline(frame, mats, np.zeros(3, np.int32), boxes, 175)
print("pose pipeline", res["kpts_img"].shape)
eng.close()
print("sanitize run ok")
