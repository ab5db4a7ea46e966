"""Seeded synthetic inputs of the BASELINE.json config shapes.

Used by bench.py, the tests and the golden-vector generator so that every
party sees the same bytes.  numpy only; nothing here touches the GPU.
"""
import numpy as np

SEED_BASE = 1234
# joints whose loss makes the reference raise (pose_estimator.py:146-157)
TORSO_JOINTS = (5, 6, 11, 12)


def heatmaps(n, J=17, H=64, W=48, seed=SEED_BASE + 1, keep_torso=True, dtype=np.float32):
    """Config 1: uniform[0,0.1) noise plus one Gaussian peak (sigma 2 px) per
    joint at a random interior location with amplitude U[0.05,1.0]; the
    shoulder/hip joints are forced >= 0.5 when keep_torso (parity variant)."""
    rng = np.random.default_rng(seed)
    hm = rng.uniform(0.0, 0.1, (n, J, H, W)).astype(np.float32)
    cy = rng.integers(3, H - 3, (n, J))
    cx = rng.integers(3, W - 3, (n, J))
    amp = rng.uniform(0.05, 1.0, (n, J)).astype(np.float32)
    if keep_torso:
        for j in TORSO_JOINTS:
            if j < J:
                amp[:, j] = np.maximum(amp[:, j], np.float32(0.5))
    yy = np.arange(H, dtype=np.float32)[None, None, :, None]
    xx = np.arange(W, dtype=np.float32)[None, None, None, :]
    g = np.exp(-((yy - cy[..., None, None]) ** 2 + (xx - cx[..., None, None]) ** 2)
               / np.float32(2 * 2.0 ** 2)).astype(np.float32)
    hm += amp[..., None, None] * g
    return hm.astype(dtype)


def person_boxes_yxyx_px(n, img_h=1080, img_w=1920, seed=SEED_BASE + 1,
                         hmin=300, hmax=1000):
    """Config 1/2 boxes: [y1,x1,y2,x2] float32 pixels inside the frame."""
    rng = np.random.default_rng(seed + 7919)
    hmax = min(hmax, img_h - 2)
    h = rng.uniform(hmin, hmax, n)
    w = np.minimum(h * rng.uniform(0.3, 0.6, n), img_w - 2)
    y1 = rng.uniform(0, img_h - h)
    x1 = rng.uniform(0, img_w - w)
    return np.stack([y1, x1, y1 + h, x1 + w], 1).astype(np.float32)


def frame_u8(h=1080, w=1920, seed=SEED_BASE + 2, smooth=True):
    """Config 2 frame: white noise, optionally blurred (three 5-tap box passes,
    sigma ~ 2.4 px) and re-stretched to the full 0..255 range."""
    rng = np.random.default_rng(seed)
    img = rng.integers(0, 256, (h, w, 3), dtype=np.uint8)
    if not smooth:
        return img
    f = img.astype(np.float32)
    for _ in range(3):                      # 3 box passes ~ Gaussian, wrap-around
        for axis in (0, 1):
            f = sum(np.roll(f, k, axis=axis) for k in range(-2, 3)) / np.float32(5)
    f = (f - f.min()) / max(float(f.max() - f.min()), 1e-6) * 255.0
    return f.astype(np.uint8)


def yolo_decoded_head(n_persons=30, n_distract=300, N=25200, nc=80, seed=SEED_BASE + 3,
                      in_size=640):
    """Config 3: decoded YOLOv5 head (1,N,5+nc) f32 with background
    obj~U[0,0.05], `n_persons` planted persons x ~30 firing anchors each (xywh
    jitter sigma 3 px, obj in [0.5,1], cls0 in [0.9,1]) and non-person
    distractors.  Returns (pred, planted_boxes_xywh)."""
    rng = np.random.default_rng(seed)
    pred = np.empty((1, N, 5 + nc), np.float32)
    pred[0, :, 0:2] = rng.uniform(0, in_size, (N, 2))
    pred[0, :, 2:4] = rng.uniform(8, 200, (N, 2))
    pred[0, :, 4] = rng.uniform(0, 0.05, N)
    pred[0, :, 5:] = rng.uniform(0, 0.2, (N, nc))
    free = rng.permutation(N)
    pos = 0
    planted = []
    for _ in range(n_persons):
        w = rng.uniform(20, 90)
        h = w * rng.uniform(2.0, 3.2)
        cx = rng.uniform(w / 2 + 2, in_size - w / 2 - 2)
        cy = rng.uniform(140 + h / 2, 500 - h / 2) if h < 350 else 320.0
        planted.append((cx, cy, w, h))
        k = int(rng.integers(25, 36))
        rows = free[pos:pos + k]
        pos += k
        pred[0, rows, 0] = cx + rng.normal(0, 3, k)
        pred[0, rows, 1] = cy + rng.normal(0, 3, k)
        pred[0, rows, 2] = w + rng.normal(0, 3, k)
        pred[0, rows, 3] = h + rng.normal(0, 3, k)
        pred[0, rows, 4] = rng.uniform(0.5, 1.0, k)
        pred[0, rows, 5] = rng.uniform(0.9, 1.0, k)
    rows = free[pos:pos + n_distract]
    pred[0, rows, 4] = rng.uniform(0.45, 1.0, len(rows))
    cls = rng.integers(1, nc, len(rows))
    pred[0, rows, 5 + cls] = rng.uniform(0.9, 1.0, len(rows))
    return pred, np.asarray(planted, np.float32)


def edet_outputs(n_frames=16, n_persons=16, img_h=1080, img_w=1920, seed=SEED_BASE + 4,
                 person_class=1.0):
    """Config 4: synthetic EfficientDet outputs per frame: boxes (F,100,4) yxyx
    px, scores (F,100) descending, classes (F,100) (person == 1.0,
    models/conv.py:22)."""
    rng = np.random.default_rng(seed)
    boxes = np.zeros((n_frames, 100, 4), np.float32)
    scores = np.zeros((n_frames, 100), np.float32)
    classes = np.zeros((n_frames, 100), np.float32)
    for f in range(n_frames):
        b = person_boxes_yxyx_px(100, img_h, img_w, seed=seed + 31 * f, hmin=200, hmax=900)
        s = np.sort(rng.uniform(0.05, 0.99, 100).astype(np.float32))[::-1]
        c = rng.integers(2, 90, 100).astype(np.float32)
        person_rows = np.sort(rng.choice(40, n_persons, replace=False))
        c[person_rows] = person_class
        s[person_rows] = np.maximum(s[person_rows], np.float32(0.75))
        boxes[f], scores[f], classes[f] = b, s, c
    return boxes, scores, classes


def yolo_head_grid(n_cols=16, n_rows=4, N=25200, nc=80, seed=SEED_BASE + 8, in_size=640, content=(140, 500)):
    """Config 2 head for a fixed person count: decoded YOLOv5 head (1,N,5+nc) f32 whose NMS keeps exactly
    n_cols*n_rows persons (64): one person per cell of a grid over the letterboxed content rows, ~12 firing anchors
    each (xywh jitter sigma 1 px), background obj~U[0,0.05], 200 non-person distractors.  Returns (pred, boxes_xywh)."""
    rng = np.random.default_rng(seed)
    pred = np.empty((1, N, 5 + nc), np.float32)
    pred[0, :, 0:2] = rng.uniform(0, in_size, (N, 2))
    pred[0, :, 2:4] = rng.uniform(8, 200, (N, 2))
    pred[0, :, 4] = rng.uniform(0, 0.05, N)
    pred[0, :, 5:] = rng.uniform(0, 0.2, (N, nc))
    free = rng.permutation(N)
    pos = 0
    cw, chh = in_size / n_cols, (content[1] - content[0]) / n_rows
    planted = []
    for r in range(n_rows):
        for c in range(n_cols):
            w = rng.uniform(0.55 * cw, 0.8 * cw)
            h = rng.uniform(0.65 * chh, 0.88 * chh)
            cx, cy = (c + 0.5) * cw, content[0] + (r + 0.5) * chh
            planted.append((cx, cy, w, h))
            k = 12
            rows = free[pos:pos + k]
            pos += k
            pred[0, rows, 0] = cx + rng.normal(0, 1, k)
            pred[0, rows, 1] = cy + rng.normal(0, 1, k)
            pred[0, rows, 2] = w + rng.normal(0, 1, k)
            pred[0, rows, 3] = h + rng.normal(0, 1, k)
            pred[0, rows, 4] = rng.uniform(0.6, 1.0, k)
            pred[0, rows, 5] = rng.uniform(0.9, 1.0, k)
    rows = free[pos:pos + 200]
    pred[0, rows, 4] = rng.uniform(0.45, 1.0, len(rows))
    cls = rng.integers(1, nc, len(rows))
    pred[0, rows, 5 + cls] = rng.uniform(0.9, 1.0, len(rows))
    return pred, np.asarray(planted, np.float32)
