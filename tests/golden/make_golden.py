#!/usr/bin/env python
"""Generate tests/golden/*.npz by EXECUTING THE REFERENCE'S OWN FUNCTIONS.

Run in the authoring container only (needs /root/reference, cv2, torchvision):

    python tests/golden/make_golden.py

What is executed, unmodified, through oracle/ref_shim.py:
  * modules/pose_estimator.py  PoseEstimator.get_max_pred_keypts_from_heatmap,
    PoseEstimator.get_keypoint_dist_dict           (decode_*.npz)
  * the per-person loop body of person_det_pose_edet4_trtserver.py:148-171 is
    DRIVEN here line for line around those two calls (the script itself cannot
    be imported: it needs tritonclient)
  * modules/onnx_utils.py  non_max_suppression (-> torchvision.ops.nms),
    w_non_max_suppression, scale_coords, letterbox_image       (nms_*.npz, misc.npz)
  * cv2.warpAffine / cv2.resize / modules/pose_estimator.py PoseEstimator.preprocess
    (crop_small.npz) -- cv2 is the north star's crop gate.
Large inputs are not stored: they are regenerated from
human_body_proportion_estimation_b200/synth.py and guarded by a sha256.
"""
import hashlib
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)

from oracle import ref_shim                      # noqa: E402
from human_body_proportion_estimation_b200 import synth   # noqa: E402

pe_mod, ou, ut = ref_shim.load()
PoseEstimator = pe_mod.PoseEstimator
import cv2        # noqa: E402
import torch      # noqa: E402

THRES = [0.45, 0.46, 0.45, 0.40, 0.34, 0.10, 0.10, 0.10, 0.10,
         0.24, 0.30, 0.11, 0.10, 0.15, 0.10, 0.25, 0.20]   # ref :62-63
KEYS = ["shoulder", "torso", "lshoulder_lelbow", "rshoulder_relbow", "lwrist_lelbow",
        "rwrist_relbow", "rhip_lhip", "rhip_rknee", "lhip_lknee", "rankle_rknee",
        "lankle_lknee"]


def sha(a):
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()


def drive_person_loop(heatmaps, boxes_px, p_height):
    """person_det_pose_edet4_trtserver.py:148-171 around the reference calls."""
    n = heatmaps.shape[0]
    xy_hm = np.zeros((n, 17, 2), np.float32)
    score = np.zeros((n, 17), np.float32)
    xy_img = np.zeros((n, 17, 2), np.float32)
    ign = np.zeros((n,), np.uint32)
    lengths = np.zeros((n, 11), np.float64)
    is_f64 = np.zeros((n, 11), bool)
    raised = np.zeros((n,), bool)
    for i, (heatmap, box) in enumerate(zip(heatmaps, boxes_px)):
        keypts, keypts_score = PoseEstimator.get_max_pred_keypts_from_heatmap(heatmap)
        xy_hm[i], score[i] = keypts, keypts_score[:, 0]
        x1, y1 = int(box[1]), int(box[0])
        x2, y2 = int(box[3]), int(box[2])
        _, hh, hw = heatmap.shape
        cw, ch = x2 - x1, y2 - y1
        keypts /= [hw, hh]
        keypts *= [cw, ch]
        keypts += [x1, y1]
        xy_img[i] = keypts
        ig = {j for j, s in enumerate(keypts_score) if s < THRES[j]}
        for j in ig:
            ign[i] |= np.uint32(1 << j)
        height_cm = p_height[min(i, len(p_height) - 1)]
        pixel_to_cm = height_cm / (y2 - y1)
        try:
            d = PoseEstimator.get_keypoint_dist_dict(pixel_to_cm, keypts, ignored_kp_idx=ig)
        except UnboundLocalError:
            raised[i] = True
            continue
        for k, key in enumerate(KEYS):
            v = d[key]
            if not isinstance(v, str):
                lengths[i, k] = float(v)
                is_f64[i, k] = isinstance(v, np.float64) or isinstance(v, float)
    return dict(xy_hm=xy_hm, score=score, xy_img=xy_img, ignored=ign,
                lengths=lengths, is_f64=is_f64, raised=raised)


def gen_decode():
    # --- small, stored in full, with the edge cases of SURVEY.md 8c
    rng = np.random.default_rng(99)
    hm = synth.heatmaps(8, 17, 16, 12, seed=5, keep_torso=True)
    hm[1, 0] = -np.abs(hm[1, 0]) - 0.5                  # all negative
    hm[1, 3, 2, 5] = np.nan                             # NaN inside a map
    hm[2, 4] = 0.0                                      # all equal -> idx 0, score 0
    hm[2, 7, :, :] = 0.25
    hm[2, 7, 10, 5] = hm[2, 7, 3, 7] = 1.0              # tie -> (7,3)
    hm[3] = synth.heatmaps(1, 17, 16, 12, seed=6, keep_torso=False)[0]
    hm[3, 5] *= 0.01                                    # shoulder lost -> ref raises
    hm[4, 9] = hm[4, 9] * 0 + np.float32(0.24)          # score == threshold (f32) -> kept
    hm[5, 13, 8, 6] = 2.0
    hm[5, 14, 8, 6] = 2.0                               # coincident knees
    boxes = synth.person_boxes_yxyx_px(8, 1080, 1920, seed=11)
    out = drive_person_loop(hm.copy(), boxes, [175, 160.5, 193])
    np.savez_compressed(os.path.join(HERE, "decode_small.npz"), heatmaps=hm, boxes_px=boxes,
                        p_height=np.array([175, 160.5, 193]), **out)
    # --- config 1 (32,17,64,48): outputs only
    hm = synth.heatmaps(32)
    boxes = synth.person_boxes_yxyx_px(32)
    out = drive_person_loop(hm.copy(), boxes, [175])
    assert not out["raised"].any()
    np.savez_compressed(os.path.join(HERE, "decode_cfg1.npz"), hm_sha=sha(hm), box_sha=sha(boxes),
                        **out)
    # --- config 1 variant where torso joints may drop (reference raises there)
    hm = synth.heatmaps(32, seed=synth.SEED_BASE + 101, keep_torso=False)
    out = drive_person_loop(hm.copy(), boxes, [175])
    np.savez_compressed(os.path.join(HERE, "decode_cfg1_drop.npz"), hm_sha=sha(hm), box_sha=sha(boxes),
                        **out)
    # --- 96x72 maps (384x288 model)
    hm = synth.heatmaps(6, 17, 96, 72, seed=77)
    boxes = synth.person_boxes_yxyx_px(6, seed=78)
    out = drive_person_loop(hm.copy(), boxes, [180, 165])
    np.savez_compressed(os.path.join(HERE, "decode_96x72.npz"), hm_sha=sha(hm), box_sha=sha(boxes),
                        **out)


def pack_list(lst, width):
    n = max([0] + [0 if x is None else x.shape[0] for x in lst])
    arr = np.zeros((len(lst), n, width), np.float32)
    cnt = np.zeros((len(lst),), np.int32)
    for i, x in enumerate(lst):
        if x is None:
            cnt[i] = -1
            continue
        x = x.numpy() if hasattr(x, "numpy") else np.asarray(x)
        arr[i, :x.shape[0]] = x
        cnt[i] = x.shape[0]
    return arr, cnt


def gen_nms():
    # --- small head, stored in full: 2 images x 600 rows x (5+8 classes)
    rng = np.random.default_rng(3)
    B, N, nc = 2, 600, 8
    pred = np.zeros((B, N, 5 + nc), np.float32)
    pred[..., 0:2] = rng.uniform(50, 590, (B, N, 2))
    pred[..., 2:4] = rng.uniform(10, 120, (B, N, 2))
    pred[..., 4] = rng.uniform(0, 1, (B, N))
    pred[..., 5:] = rng.uniform(0, 1, (B, N, nc))
    # clusters around a few centres so suppression really happens
    for b in range(B):
        for c in range(12):
            rows = rng.choice(N, 20, replace=False)
            ctr = rng.uniform(100, 540, 2)
            wh = rng.uniform(40, 120, 2)
            pred[b, rows, 0:2] = ctr + rng.normal(0, 4, (20, 2))
            pred[b, rows, 2:4] = wh + rng.normal(0, 4, (20, 2))
            pred[b, rows, 4] = rng.uniform(0.5, 1, 20)
            pred[b, rows, 5 + c % nc] = rng.uniform(0.85, 1, 20)
    # threshold edges: obj exactly at float32(0.4) and one ulp above
    pred[0, 0, 4] = np.float32(0.4)
    pred[0, 1, 4] = np.nextafter(np.float32(0.4), np.float32(1))
    pred[0, 1, 5:] = 1.0
    # equal scores (stable order) and an exact duplicate
    pred[1, 10] = pred[1, 11] = pred[1, 12]
    pred[1, 10:13, 0] += np.array([0, 300, 0], np.float32)
    t = torch.from_numpy(pred.copy())
    o1 = ou.non_max_suppression(t.clone(), conf_thres=0.4, iou_thres=0.5)
    o2 = ou.non_max_suppression(t.clone(), conf_thres=0.4, iou_thres=0.5, classes=[0, 3])
    o3 = ou.non_max_suppression(t.clone(), conf_thres=0.25, iou_thres=0.45)
    leg_in = t.clone()
    o4 = ou.w_non_max_suppression(leg_in, nc, conf_thres=0.4, nms_thres=0.3)
    a1, c1 = pack_list(o1, 6)
    a2, c2 = pack_list(o2, 6)
    a3, c3 = pack_list(o3, 6)
    a4, c4 = pack_list(o4, 7)
    np.savez_compressed(os.path.join(HERE, "nms_small.npz"), pred=pred,
                        off_a=a1, off_a_n=c1, off_b=a2, off_b_n=c2, off_c=a3, off_c_n=c3,
                        leg=a4, leg_n=c4, leg_mutated=leg_in.numpy())
    # --- known-answer boxes straight through torchvision.ops.nms (reference call site :205)
    import torchvision
    boxes = np.array([[0, 0, 10, 10], [0, 0, 10, 5],      # IoU exactly 0.5 -> both kept
                      [20, 20, 30, 30], [20, 20, 30, 30],  # duplicate -> second dropped
                      [40, 40, 40, 40], [40, 40, 40, 40],  # zero area -> NaN -> both kept
                      [50, 50, 60, 60], [51, 51, 61, 61]], np.float32)
    scores = np.array([0.9, 0.8, 0.7, 0.7, 0.6, 0.6, 0.5, 0.5], np.float32)
    keep = torchvision.ops.nms(torch.from_numpy(boxes), torch.from_numpy(scores), 0.5).numpy()
    rb = rng.uniform(0, 600, (2000, 2)).astype(np.float32)
    rwh = rng.uniform(5, 150, (2000, 2)).astype(np.float32)
    rboxes = np.concatenate([rb, rb + rwh], 1)
    rscores = np.round(rng.uniform(0, 1, 2000), 2).astype(np.float32)     # many ties
    keeps = {}
    for thr in (0.3, 0.45, 0.5, 0.7):
        keeps["rand_keep_%d" % int(thr * 100)] = torchvision.ops.nms(
            torch.from_numpy(rboxes), torch.from_numpy(rscores), thr).numpy()
    np.savez_compressed(os.path.join(HERE, "nms_kat.npz"), boxes=boxes, scores=scores, keep=keep,
                        rand_boxes=rboxes, rand_scores=rscores, **keeps)
    # --- config 3 head: outputs only
    pred, _ = synth.yolo_decoded_head()
    t = torch.from_numpy(pred.copy())
    oa = ou.non_max_suppression(t.clone(), conf_thres=0.4, iou_thres=0.5)
    ob = ou.non_max_suppression(t.clone(), conf_thres=0.4, iou_thres=0.5, classes=[0])
    aa, ca = pack_list(oa, 6)
    ab, cb = pack_list(ob, 6)
    np.savez_compressed(os.path.join(HERE, "nms_cfg3.npz"), pred_sha=sha(pred),
                        all_cls=aa, all_cls_n=ca, person=ab, person_n=cb)


def gen_crop():
    rng = np.random.default_rng(8)
    img = rng.integers(0, 256, (120, 160, 3), dtype=np.uint8)       # white noise: worst case
    oh, ow = 64, 48
    mats = []
    for t in range(10):
        sx, sy = rng.uniform(0.4, 2.5, 2)
        ang = 0.0 if t < 6 else rng.uniform(-0.6, 0.6)
        mats.append([[sx * np.cos(ang), -sy * np.sin(ang), rng.uniform(-15, 90)],
                     [sx * np.sin(ang), sy * np.cos(ang), rng.uniform(-15, 60)]])
    mats = np.asarray(mats, np.float64)
    f32 = np.stack([cv2.warpAffine(img.astype(np.float32), M, (ow, oh),
                                   flags=cv2.INTER_LINEAR | cv2.WARP_INVERSE_MAP,
                                   borderMode=cv2.BORDER_CONSTANT, borderValue=0) for M in mats])
    u8 = np.stack([cv2.warpAffine(img, M, (ow, oh), flags=cv2.INTER_LINEAR | cv2.WARP_INVERSE_MAP,
                                  borderMode=cv2.BORDER_CONSTANT, borderValue=0) for M in mats])
    rs = cv2.resize(img, (72, 96))
    rs2 = cv2.resize(img, (333, 201))
    pre = PoseEstimator.preprocess(img[None].copy(), w=72, h=96)     # ref pose_estimator.py:29-45
    np.savez_compressed(os.path.join(HERE, "crop_small.npz"), img=img, mats=mats, warp_f32=f32,
                        warp_u8=u8, resize_72x96=rs, resize_333x201=rs2, hrnet_pre_72x96=pre)


def gen_misc():
    from PIL import Image
    c = np.array([[100, 200, 300, 400], [0, 139, 640, 501], [-5, 10, 700, 650]], np.float32)
    sc = ou.scale_coords((640, 640), c.copy(), (1080, 1920))
    sc2 = ou.scale_coords((640, 640), c.copy(), (2160, 3840))
    lb = ou.letterbox_image(Image.new("RGB", (1920, 1080), (10, 20, 30)), (640, 640))
    lb = np.asarray(lb)
    rows = np.nonzero((lb[:, 320] != 128).any(-1))[0]
    np.savez_compressed(os.path.join(HERE, "misc.npz"), coords=c, scaled_1080=sc, scaled_2160=sc2,
                        letterbox_rows=np.array([rows.min(), rows.max()]),
                        letterbox_pad=lb[0, 0])


def gen_letterbox_pil():
    """reference modules/onnx_utils.py:225-235 letterbox_image (PIL BICUBIC, antialiased) executed as is."""
    import hashlib
    from PIL import Image
    import PIL
    sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
    from human_body_proportion_estimation_b200 import synth
    rng = np.random.default_rng(77)
    out = {"pil_version": np.array(PIL.__version__)}
    cases = [((97, 131), (64, 48)), ((20, 30), (64, 64)), ((108, 192), (64, 64)), ((64, 64), (64, 64)), ((33, 250), (96, 40))]
    for i, ((ih, iw), (w, h)) in enumerate(cases):
        img = rng.integers(0, 256, (ih, iw, 3), dtype=np.uint8)
        lb = np.asarray(ou.letterbox_image(Image.fromarray(img), (w, h)))
        out["img%d" % i] = img
        out["size%d" % i] = np.array([w, h])
        out["lb%d" % i] = lb
    # config 3: the synthetic 1080p frame (white noise = worst case for a resampler) -> 640 x 640
    frame = synth.frame_u8(smooth=False)
    lb = np.asarray(ou.letterbox_image(Image.fromarray(frame), (640, 640)))
    out["lb1080_sha256"] = np.array(hashlib.sha256(lb.tobytes()).hexdigest())
    out["lb1080_sample"] = lb[::8, ::8].copy()
    np.savez_compressed(os.path.join(HERE, "letterbox_pil.npz"), **out)


if __name__ == "__main__":
    gen_decode()
    gen_nms()
    gen_crop()
    gen_misc()
    gen_letterbox_pil()
    for f in sorted(os.listdir(HERE)):
        if f.endswith(".npz"):
            print(f, os.path.getsize(os.path.join(HERE, f)))
