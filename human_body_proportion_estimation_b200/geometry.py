"""Host-side crop geometry: per-person 2x3 dst->src matrices for hbp_crop_warp."""
import numpy as np


def crop_and_resize_matrices(boxes_yxyx_norm, img_h, img_w, out_h, out_w):
    """tf.image.crop_and_resize mapping used by the reference's ensemble
    (models/conv.py:61-70, align-corners): dst (i,j) -> src
    y = y1*(H-1) + i*(y2-y1)*(H-1)/(out_h-1), x likewise.  -> (P,2,3) float64"""
    b = np.asarray(boxes_yxyx_norm, np.float64).reshape(-1, 4)
    M = np.zeros((b.shape[0], 2, 3), np.float64)
    M[:, 0, 0] = (b[:, 3] - b[:, 1]) * (img_w - 1) / max(out_w - 1, 1)
    M[:, 0, 2] = b[:, 1] * (img_w - 1)
    M[:, 1, 1] = (b[:, 2] - b[:, 0]) * (img_h - 1) / max(out_h - 1, 1)
    M[:, 1, 2] = b[:, 0] * (img_h - 1)
    return M


def box_resize_matrices(boxes_xyxy_px, out_h, out_w):
    """cv2.resize-style (half-pixel) stretch of a pixel box onto the crop
    (modules/pose_estimator.py:41 applied to a box).  -> (P,2,3) float64"""
    b = np.asarray(boxes_xyxy_px, np.float64).reshape(-1, 4)
    M = np.zeros((b.shape[0], 2, 3), np.float64)
    sx, sy = (b[:, 2] - b[:, 0]) / out_w, (b[:, 3] - b[:, 1]) / out_h
    M[:, 0, 0], M[:, 0, 2] = sx, b[:, 0] + 0.5 * sx - 0.5
    M[:, 1, 1], M[:, 1, 2] = sy, b[:, 1] + 0.5 * sy - 0.5
    return M
