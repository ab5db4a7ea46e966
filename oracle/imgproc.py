"""Oracle (TEST INFRASTRUCTURE): image resampling restated in numpy.

Reference anchors (relative to /root/reference):
  * per-person crop ........ models/conv.py:59-80 (tf.image.crop_and_resize,
      align-corners bilinear of the /255 image) -- TensorFlow is absent, so the
      crop GEOMETRY is restated (`crop_and_resize_matrix`) and the SAMPLING gate
      is the north star's: cv2.warpAffine(INTER_LINEAR, WARP_INVERSE_MAP,
      BORDER_CONSTANT 0), whose fixed-point arithmetic `warp_affine_cv2` below
      reproduces bit for bit (pinned by tests/golden/crop_*.npz, generated with
      cv2 4.13 in the authoring container).
  * HRNet preprocess ....... human_body_length_est/modules/pose_estimator.py:29-45
      (cv2.cvtColor BGR2RGB, cv2.resize u8 bilinear, /255.0, CHW float32);
      `resize_linear_u8_cv2` restates cv2.resize's 11-bit fixed point.
  * edet preprocess ........ human_body_length_est/person_det_pose_edet4_trtserver.py:15-18
  * YOLO letterbox ......... human_body_length_est/modules/onnx_utils.py:225-235 +
      obj_det_yolov5_onnx.py:27-36 (PIL BICUBIC in the reference; the bilinear
      letterbox here is this build's documented stand-in).
"""
import numpy as np

AB_BITS = 10            # cv2 warpAffine: coordinates carried in 1/1024 px
INTER_BITS = 5          # ... and quantised to 1/32 px for the bilinear weights
INTER_TAB = 1 << INTER_BITS


def warp_affine_coords(M, out_h, out_w):
    """cv2 imgwarp.cpp WarpAffineInvoker: fixed-point source coordinates for
    every destination pixel.  M is the 2x3 dst->src matrix (float64).  Returns
    integer (sx, sy) and the 1/32 fractions (fx, fy) as int32 arrays."""
    M = np.asarray(M, np.float64).reshape(2, 3)
    scale = float(1 << AB_BITS)
    xs = np.arange(out_w, dtype=np.float64)
    ys = np.arange(out_h, dtype=np.float64)
    adelta = np.rint(M[0, 0] * xs * scale).astype(np.int64)
    bdelta = np.rint(M[1, 0] * xs * scale).astype(np.int64)
    half = (1 << AB_BITS) // INTER_TAB // 2            # 16
    X0 = np.rint((M[0, 1] * ys + M[0, 2]) * scale).astype(np.int64) + half
    Y0 = np.rint((M[1, 1] * ys + M[1, 2]) * scale).astype(np.int64) + half
    X = (X0[:, None] + adelta[None, :]) >> (AB_BITS - INTER_BITS)
    Y = (Y0[:, None] + bdelta[None, :]) >> (AB_BITS - INTER_BITS)
    return (X >> INTER_BITS).astype(np.int64), (Y >> INTER_BITS).astype(np.int64), \
        (X & (INTER_TAB - 1)).astype(np.int32), (Y & (INTER_TAB - 1)).astype(np.int32)


def warp_affine_cv2(src, M, out_h, out_w):
    """cv2.warpAffine(src.astype(f32), M, (out_w,out_h), INTER_LINEAR |
    WARP_INVERSE_MAP, BORDER_CONSTANT, 0) -> (out_h,out_w,C) float32.

    Each tap is the source value or 0 outside the image; weights are products of
    the two 1/32 tables in float32; accumulation order is cv2's remapBilinear:
    S00*w00 + S01*w01 + S10*w10 + S11*w11 (left to right, float32)."""
    src = np.asarray(src)
    H, W = src.shape[:2]
    C = 1 if src.ndim == 2 else src.shape[2]
    s = src.reshape(H, W, C).astype(np.float32)
    sx, sy, fx, fy = warp_affine_coords(M, out_h, out_w)
    tab = (np.arange(INTER_TAB, dtype=np.float32) / np.float32(INTER_TAB))
    ax1, ay1 = tab[fx], tab[fy]
    ax0, ay0 = np.float32(1) - ax1, np.float32(1) - ay1
    w00, w01, w10, w11 = ay0 * ax0, ay0 * ax1, ay1 * ax0, ay1 * ax1

    def tap(yy, xx):
        ok = (yy >= 0) & (yy < H) & (xx >= 0) & (xx < W)
        v = s[np.clip(yy, 0, H - 1), np.clip(xx, 0, W - 1)]
        return np.where(ok[..., None], v, np.float32(0))

    out = tap(sy, sx) * w00[..., None]
    out = out + tap(sy, sx + 1) * w01[..., None]
    out = out + tap(sy + 1, sx) * w10[..., None]
    out = out + tap(sy + 1, sx + 1) * w11[..., None]
    return out.astype(np.float32)


def crop_and_resize_matrix(box_yxyx_norm, img_h, img_w, out_h, out_w):
    """models/conv.py:61-70 geometry (tf.image.crop_and_resize, align corners):
    dst (i,j) -> src y = y1*(H-1) + i*(y2-y1)*(H-1)/(out_h-1), x likewise.
    Returned as the 2x3 dst->src matrix (float64) fed to the crop kernel."""
    y1, x1, y2, x2 = (float(v) for v in box_yxyx_norm)
    sy = (y2 - y1) * (img_h - 1) / (out_h - 1) if out_h > 1 else 0.0
    sx = (x2 - x1) * (img_w - 1) / (out_w - 1) if out_w > 1 else 0.0
    return np.array([[sx, 0.0, x1 * (img_w - 1)],
                     [0.0, sy, y1 * (img_h - 1)]], np.float64)


def box_resize_matrix(box_xyxy_px, out_h, out_w):
    """cv2.resize-style half-pixel mapping of a pixel box onto the crop
    (pose_estimator.py:41 applied to frame[y1:y2, x1:x2])."""
    x1, y1, x2, y2 = (float(v) for v in box_xyxy_px)
    sx, sy = (x2 - x1) / out_w, (y2 - y1) / out_h
    return np.array([[sx, 0.0, x1 + 0.5 * sx - 0.5],
                     [0.0, sy, y1 + 0.5 * sy - 0.5]], np.float64)


def crop_persons(frame_u8, mats, out_h, out_w, swap_rb=True, out_dtype=np.float16):
    """The crop stage as this build defines it: cv2-exact bilinear of the u8
    frame, optional channel swap, /255 in float32, NCHW, cast to out_dtype."""
    crops = []
    for M in mats:
        c = warp_affine_cv2(frame_u8, M, out_h, out_w)
        if swap_rb:
            c = c[..., ::-1]
        c = (c / np.float32(255.0)).astype(np.float32)
        crops.append(np.transpose(c, (2, 0, 1)).astype(out_dtype))
    return np.stack(crops) if crops else np.zeros((0, 3, out_h, out_w), out_dtype)


# --------------------------------------------------------------------------
# cv2.resize(u8, INTER_LINEAR): 11-bit fixed-point separable bilinear
# --------------------------------------------------------------------------
RESIZE_COEF_BITS = 11
RESIZE_COEF_SCALE = 1 << RESIZE_COEF_BITS


def _resize_axis_table(dst_n, src_n, vertical=False):
    """cv2 resize.cpp: per-destination index, source index pair and the two
    int16 coefficients (saturate_cast<short>(w * 2048), round-half-even).
    Horizontally cv2 zeroes the fraction when the left tap is clamped; vertically
    it only clamps the two ROW INDICES and keeps the fraction, so the first/last
    output rows blend a source row with itself (and lose up to one level to the
    two truncations)."""
    scale = 1.0 / (dst_n / src_n)
    d = np.arange(dst_n, dtype=np.float64)
    f = ((d + 0.5) * scale - 0.5).astype(np.float32)        # float in cv2
    s0 = np.floor(f).astype(np.int64)
    frac = (f - s0.astype(np.float32)).astype(np.float32)
    if vertical:
        s1 = np.clip(s0 + 1, 0, src_n - 1)
        s0 = np.clip(s0, 0, src_n - 1)
    else:
        lo = s0 < 0
        frac[lo] = 0
        s0[lo] = 0
        hi = s0 >= src_n - 1
        frac[hi] = 0
        s0[hi] = src_n - 1
        s1 = np.minimum(s0 + 1, src_n - 1)
    c0 = np.rint((np.float32(1) - frac) * np.float32(RESIZE_COEF_SCALE)).astype(np.int64)
    c1 = np.rint(frac * np.float32(RESIZE_COEF_SCALE)).astype(np.int64)
    return s0, s1, c0, c1


def resize_linear_u8_cv2(src, out_w, out_h):
    """cv2.resize(src_u8, (out_w,out_h)) for INTER_LINEAR, any channel count."""
    src = np.asarray(src, np.uint8)
    H, W = src.shape[:2]
    s = src.reshape(H, W, -1).astype(np.int64)
    x0, x1, a0, a1 = _resize_axis_table(out_w, W)
    y0, y1, b0, b1 = _resize_axis_table(out_h, H, vertical=True)
    rows = s[:, x0] * a0[None, :, None] + s[:, x1] * a1[None, :, None]   # (H,out_w,C)
    r0, r1 = rows[y0], rows[y1]
    v = (((b0[:, None, None] * (r0 >> 4)) >> 16) + ((b1[:, None, None] * (r1 >> 4)) >> 16) + 2) >> 2
    return np.clip(v, 0, 255).astype(np.uint8).reshape((out_h, out_w) + src.shape[2:])


def hrnet_preprocess(frames_bgr_u8, w=288, h=384):
    """pose_estimator.py:29-45 for a (B,H,W,3) u8 BGR array: channel swap,
    cv2.resize to (w,h) (aspect not kept), /255.0 in float64, CHW, float32."""
    out = []
    for fr in frames_bgr_u8:
        rgb = fr[..., ::-1]
        r = resize_linear_u8_cv2(rgb, w, h)
        out.append(np.transpose(r / 255.0, (2, 0, 1)).astype(np.float32))
    return np.stack(out)


def edet_preprocess(frame_bgr_u8):
    """person_det_pose_edet4_trtserver.py:15-18 with the shipped dynamic-shape
    model (width=height=None -> no resize): BGR->RGB, uint8."""
    return np.ascontiguousarray(frame_bgr_u8[..., ::-1]).astype(np.uint8)


def letterbox_linear(frame_rgb_u8, w=640, h=640, pad=128):
    """Letterbox with the reference's geometry (onnx_utils.py:225-235: scale,
    int() sizes, centred paste on grey 128) and cv2.resize bilinear as the
    resampler (the reference uses PIL BICUBIC -- see module docstring).
    Returns (3,h,w) float32 in [0,1] like obj_det_yolov5_onnx.py:33-35."""
    ih, iw = frame_rgb_u8.shape[:2]
    scale = min(w / iw, h / ih)
    nw, nh = int(iw * scale), int(ih * scale)
    canvas = np.full((h, w, 3), pad, np.uint8)
    ox, oy = (w - nw) // 2, (h - nh) // 2
    canvas[oy:oy + nh, ox:ox + nw] = resize_linear_u8_cv2(frame_rgb_u8, nw, nh)
    out = np.transpose(canvas, (2, 0, 1)).astype(np.float32)
    out /= 255.0
    return out


# ---------------------------------------------------------------------------
# PIL (Pillow) antialiased bicubic resize -- the resampler of the reference's YOLO letterbox
# (modules/onnx_utils.py:232, `image.resize((nw, nh), Image.BICUBIC)`; pillow==10.3.0 in
# requirements.txt:6).  Pillow is a third-party dependency, not under /root/reference: this restates
# its published algorithm (src/libImaging/Resample.c: precompute_coeffs, normalize_coeffs_8bpc,
# ImagingResampleHorizontal_8bpc / Vertical_8bpc) and is pinned bit for bit against the Pillow installed
# in this container (12.2.0; the 8-bit resample path is unchanged since 7.0) through
# tests/golden/letterbox_pil.npz.
# ---------------------------------------------------------------------------
PIL_PRECISION_BITS = 32 - 8 - 2


def _pil_bicubic_filter(x):
    a = -0.5
    if x < 0.0:
        x = -x
    if x < 1.0:
        return ((a + 2.0) * x - (a + 3.0)) * x * x + 1
    if x < 2.0:
        return (((x - 5) * x + 8) * x - 4) * a
    return 0.0


def pil_bicubic_coeffs(in_size, out_size):
    """-> (ksize, bounds[out_size,2] = (first tap, tap count), kk[out_size,ksize] int32 fixed point)."""
    scale = filterscale = float(in_size) / out_size
    if filterscale < 1.0:
        filterscale = 1.0
    support = 2.0 * filterscale
    ksize = int(np.ceil(support)) * 2 + 1
    bounds = np.zeros((out_size, 2), np.int32)
    kk = np.zeros((out_size, ksize), np.int32)
    ss = 1.0 / filterscale
    for xx in range(out_size):
        center = 0.0 + (xx + 0.5) * scale
        xmin = int(center - support + 0.5)
        if xmin < 0:
            xmin = 0
        xmax = int(center + support + 0.5)
        if xmax > in_size:
            xmax = in_size
        xmax -= xmin
        w = [_pil_bicubic_filter((x + xmin - center + 0.5) * ss) for x in range(xmax)]
        ww = 0.0
        for v in w:
            ww += v
        for x in range(xmax):
            v = w[x] / ww if ww != 0.0 else w[x]
            kk[xx, x] = int(-0.5 + v * (1 << PIL_PRECISION_BITS)) if v < 0 else int(0.5 + v * (1 << PIL_PRECISION_BITS))
        bounds[xx] = (xmin, xmax)
    return ksize, bounds, kk


def _pil_pass(img, out_size, axis):
    """one 8-bit resample pass along `axis` (0 = vertical, 1 = horizontal) of an (H,W,C) uint8 image"""
    src = np.moveaxis(img, axis, 0).astype(np.int64)
    _, bounds, kk = pil_bicubic_coeffs(src.shape[0], out_size)
    out = np.empty((out_size,) + src.shape[1:], np.uint8)
    for xx in range(out_size):
        lo, n = bounds[xx]
        acc = np.tensordot(kk[xx, :n].astype(np.int64), src[lo:lo + n], axes=(0, 0)) + (1 << (PIL_PRECISION_BITS - 1))
        out[xx] = np.clip(acc >> PIL_PRECISION_BITS, 0, 255)
    return np.moveaxis(out, 0, axis)


def resize_bicubic_pil(img_u8, w, h):
    """PIL.Image.resize((w, h), BICUBIC) of an (H,W,3) uint8 array: horizontal pass into a uint8
    intermediate, then the vertical pass (Resample.c ImagingResampleInner)."""
    ih, iw = img_u8.shape[:2]
    out = img_u8
    if w != iw:
        out = _pil_pass(out, w, 1)
    if h != ih:
        out = _pil_pass(out, h, 0)
    return np.ascontiguousarray(out)


def letterbox_pil(frame_rgb_u8, w=640, h=640, pad=128):
    """The reference's letterbox + preprocess (modules/onnx_utils.py:225-235,
    obj_det_yolov5_onnx.py:27-36): PIL-bicubic resize to int(iw*scale) x int(ih*scale), centred
    paste on grey 128, CHW float32 / 255."""
    ih, iw = frame_rgb_u8.shape[:2]
    scale = min(w / iw, h / ih)
    nw, nh = int(iw * scale), int(ih * scale)
    canvas = np.full((h, w, 3), pad, np.uint8)
    ox, oy = (w - nw) // 2, (h - nh) // 2
    canvas[oy:oy + nh, ox:ox + nw] = resize_bicubic_pil(frame_rgb_u8, nw, nh)
    out = np.transpose(canvas, (2, 0, 1)).astype(np.float32)
    out /= 255.0
    return out
