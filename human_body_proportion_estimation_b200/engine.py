"""Per-GPU engine: a thin, numpy-facing layer over the C ABI (include/hbp.h).

One `Engine` = one hbp_ctx = one GPU; it is the in-process replacement of the
reference's Triton client (human_body_length_est/modules/triton_utils.py:11-34,
131-177).  Host arrays in, host arrays out; every method is one C-ABI call.
`MultiGpuEngine` shards frames across several engines with one Python thread
per GPU (ctypes drops the GIL) -- frames are independent, so there is no
collective anywhere on the path.
"""
import ctypes as C
import threading

import numpy as np

from . import _capi
from ._capi import DEVICE, F16, F32, HOST, NCHW, NHWC, U8, check, ptr

# person_det_pose_edet4_trtserver.py:62-63
KEYPOINT_THRES_LIST = (0.45, 0.46, 0.45, 0.40, 0.34, 0.10, 0.10, 0.10, 0.10,
                       0.24, 0.30, 0.11, 0.10, 0.15, 0.10, 0.25, 0.20)
SEGMENT_KEYS = ("shoulder", "torso", "lshoulder_lelbow", "rshoulder_relbow", "lwrist_lelbow",
                "rwrist_relbow", "rhip_lhip", "rhip_rknee", "lhip_lknee", "rankle_rknee",
                "lankle_lknee")
NOT_VISIBLE = "Part not visible"
_NP2HBP = {np.dtype(np.uint8): U8, np.dtype(np.float16): F16, np.dtype(np.float32): F32}


def _c(a, dtype):
    return np.ascontiguousarray(a, dtype=dtype)


class Engine:
    def __init__(self, device=0):
        self._lib = _capi.lib()
        self._ctx = C.c_void_p()
        check(self._lib.hbp_ctx_create(int(device), C.byref(self._ctx)))
        self.device = int(device)
        self.hrnet = None           # (width, in_h, in_w) once loaded

    def close(self):
        if self._ctx:
            self._lib.hbp_ctx_destroy(self._ctx)
            self._ctx = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # ---- plumbing --------------------------------------------------------
    def sync(self):
        check(self._lib.hbp_sync(self._ctx))

    def dev_alloc(self, nbytes):
        p = C.c_void_p()
        check(self._lib.hbp_dev_alloc(self._ctx, int(nbytes), C.byref(p)))
        return p.value

    def dev_free(self, p):
        check(self._lib.hbp_dev_free(self._ctx, C.c_void_p(p)))

    def pinned_empty(self, shape, dtype):
        """numpy array backed by pinned host memory (freed with the engine)."""
        dtype = np.dtype(dtype)
        n = int(np.prod(shape)) * dtype.itemsize
        p = C.c_void_p()
        check(self._lib.hbp_host_alloc(self._ctx, max(n, 1), C.byref(p)))
        buf = (C.c_uint8 * max(n, 1)).from_address(p.value)
        return np.frombuffer(buf, dtype=dtype, count=int(np.prod(shape))).reshape(shape)

    def h2d(self, dev_ptr, arr):
        check(self._lib.hbp_copy_h2d(self._ctx, C.c_void_p(dev_ptr), ptr(arr), arr.nbytes))

    def d2h(self, arr, dev_ptr):
        check(self._lib.hbp_copy_d2h(self._ctx, ptr(arr), C.c_void_p(dev_ptr), arr.nbytes))

    def to_device(self, arr):
        arr = np.ascontiguousarray(arr)
        p = self.dev_alloc(arr.nbytes)
        self.h2d(p, arr)
        self.sync()
        return p

    def timer_start(self, slot=0):
        check(self._lib.hbp_timer_start(self._ctx, slot))

    def timer_stop(self, slot=0):
        check(self._lib.hbp_timer_stop(self._ctx, slot))

    def timer_ms(self, slot=0):
        ms = C.c_float()
        check(self._lib.hbp_timer_elapsed_ms(self._ctx, slot, C.byref(ms)))
        return ms.value

    def flush_l2(self):
        check(self._lib.hbp_flush_l2(self._ctx))

    def kernel_launches(self):
        n = C.c_uint64()
        check(self._lib.hbp_kernel_launches(self._ctx, C.byref(n)))
        return n.value

    # ---- K1 ----------------------------------------------------------------
    def preprocess(self, frames, mode, out_h=None, out_w=None, swap_rb=True, pad_value=128,
                   out_dtype=np.float32, layout=NCHW):
        frames = _c(frames, np.uint8)
        if frames.ndim == 3:
            frames = frames[None]
        n, h, w, _ = frames.shape
        out_h, out_w = out_h or h, out_w or w
        shape = (n, 3, out_h, out_w) if layout == NCHW else (n, out_h, out_w, 3)
        out = np.empty(shape, out_dtype)
        check(self._lib.hbp_preprocess(self._ctx, ptr(frames), n, h, w, mode, out_h, out_w, int(swap_rb),
                                       pad_value, ptr(out), _NP2HBP[np.dtype(out_dtype)], layout, HOST))
        return out

    # ---- K2 / K3 -------------------------------------------------------------
    def yolo_decode_raw(self, heads, in_w=640, in_h=640):
        heads = [_c(h, np.float32) for h in heads]
        B, nc = heads[0].shape[0], heads[0].shape[-1] - 5
        S = [h.shape[2] for h in heads]
        out = np.empty((B, 3 * sum(s * s for s in S), 5 + nc), np.float32)
        check(self._lib.hbp_yolo_decode_raw(self._ctx, ptr(heads[0]), ptr(heads[1]), ptr(heads[2]), B,
                                            S[0], S[1], S[2], nc, in_w, in_h, ptr(out), HOST))
        return out

    def yolo_nms(self, pred, conf_thres=0.25, iou_thres=0.45, classes=None, max_det=300):
        pred = _c(pred, np.float32)
        B, N, E = pred.shape
        cls = _c(classes, np.int32) if classes is not None else None
        det = np.zeros((B, max_det, 6), np.float32)
        cnt = np.zeros((B,), np.int32)
        check(self._lib.hbp_yolo_nms(self._ctx, ptr(pred), B, N, E - 5, float(conf_thres), float(iou_thres),
                                     ptr(cls), 0 if cls is None else len(cls), max_det, ptr(det), ptr(cnt), HOST))
        return [det[b, :cnt[b]].copy() for b in range(B)]

    def yolo_nms_legacy(self, pred, num_classes, conf_thres=0.5, nms_thres=0.4, max_out=None):
        pred = _c(pred, np.float32)
        B, N, E = pred.shape
        num_classes = int(num_classes)
        # the reference slices the class scores [:, 5:5+num_classes] (onnx_utils.py:59): rows may be wider than
        # 5+num_classes (extra columns ignored) but never narrower; the C side strides rows by 5+nc
        if num_classes < 1 or E < 5 + num_classes:
            raise ValueError("prediction rows have %d columns, need at least 5 + num_classes = %d" % (E, 5 + num_classes))
        if E > 5 + num_classes:
            pred = np.ascontiguousarray(pred[..., :5 + num_classes])
        max_out = max_out or N
        det = np.zeros((B, max_out, 7), np.float32)
        cnt = np.zeros((B,), np.int32)
        check(self._lib.hbp_yolo_nms_legacy(self._ctx, ptr(pred), B, N, num_classes, float(conf_thres),
                                            float(nms_thres), max_out, ptr(det), ptr(cnt), HOST))
        return [None if cnt[b] < 0 else det[b, :cnt[b]].copy() for b in range(B)]

    def scale_coords(self, img1_shape, coords, img0_shape):
        """in place on a float32 (n,>=4) array like the reference (onnx_utils.py:252-266)"""
        box = _c(coords[:, :4], np.float32)
        check(self._lib.hbp_scale_coords(self._ctx, ptr(box), box.shape[0], int(img1_shape[0]),
                                         int(img1_shape[1]), int(img0_shape[0]), int(img0_shape[1]), HOST))
        coords[:, :4] = box
        return coords

    def edet_person_filter(self, boxes, scores, classes, det_thres, x_expand, y_expand, img_h, img_w,
                           max_persons=3, person_class=1.0):
        boxes, scores, classes = _c(boxes, np.float32), _c(scores, np.float32), _c(classes, np.float32)
        if boxes.ndim == 2:
            boxes, scores, classes = boxes[None], scores[None], classes[None]
        F_, K = scores.shape
        out = np.zeros((F_, max_persons, 4), np.float32)
        cnt = np.zeros((F_,), np.int32)
        check(self._lib.hbp_edet_person_filter(self._ctx, ptr(boxes), ptr(scores), ptr(classes), F_, K,
                                               float(person_class), float(det_thres), float(x_expand),
                                               float(y_expand), img_h, img_w, max_persons, ptr(out), ptr(cnt), HOST))
        return [out[f, :cnt[f]].copy() for f in range(F_)]

    # ---- K4 ----------------------------------------------------------------
    def crop_warp(self, frames, mats, frame_idx, out_h, out_w, swap_rb=True, out_dtype=np.float16):
        frames = _c(frames, np.uint8)
        if frames.ndim == 3:
            frames = frames[None]
        mats = _c(mats, np.float64).reshape(-1, 6)
        fi = _c(frame_idx, np.int32)
        P = mats.shape[0]
        out = np.empty((P, 3, out_h, out_w), out_dtype)
        check(self._lib.hbp_crop_warp(self._ctx, ptr(frames), frames.shape[0], frames.shape[1],
                                      frames.shape[2], ptr(mats), ptr(fi), P, out_h, out_w, int(swap_rb),
                                      ptr(out), _NP2HBP[np.dtype(out_dtype)], HOST))
        return out

    # ---- K5 ----------------------------------------------------------------
    def load_hrnet(self, weights=None, width=32, in_h=256, in_w=192, seed=0):
        from . import hrnet_arch
        if weights is None:
            weights = hrnet_arch.random_weights(width, in_h, in_w, seed)
        wb, bb = hrnet_arch.pack(weights, width, in_h, in_w)
        check(self._lib.hbp_hrnet_load(self._ctx, width, in_h, in_w, ptr(wb), wb.size, ptr(bb), bb.size))
        self.hrnet = (width, in_h, in_w)
        return weights

    def conv2d_nhwc(self, x, w, bias, residual=None, stride=1, up=1, relu=False, engine=1, time_iters=0):
        """x (P,H,W,Cin) f16, w (Cout,Cin,k,k) -> (P,H/stride*up,W/stride*up,Cout) f16, used_engine
        (+ average device ms per launch when time_iters > 0)"""
        x = _c(x, np.float16)
        P, H, W, Cin = x.shape
        Cout, _, k, _ = w.shape
        wb = _c(np.transpose(np.asarray(w), (2, 3, 0, 1)), np.float16)
        b = _c(bias, np.float32)
        out = np.empty((P, H // stride * up, W // stride * up, Cout), np.float16)
        r = _c(residual, np.float16) if residual is not None else None
        used = C.c_int(-1)
        if time_iters > 0:
            ms = C.c_float()
            check(self._lib.hbp_conv2d_nhwc_timed(self._ctx, int(engine), ptr(x), P, H, W, Cin, ptr(wb), ptr(b), ptr(r),
                                                  Cout, k, stride, up, int(relu), ptr(out), C.byref(used), HOST,
                                                  int(time_iters), C.byref(ms)))
            return out, used.value, ms.value
        check(self._lib.hbp_conv2d_nhwc(self._ctx, int(engine), ptr(x), P, H, W, Cin, ptr(wb), ptr(b), ptr(r),
                                        Cout, k, stride, up, int(relu), ptr(out), C.byref(used), HOST))
        return out, used.value

    def set_hrnet_engine(self, engine):
        check(self._lib.hbp_hrnet_set_engine(self._ctx, int(engine)))

    def hrnet_forward(self, crops, out_dtype=np.float32):
        crops = _c(crops, np.float16)
        P = crops.shape[0]
        _, ih, iw = self.hrnet
        assert crops.shape[1:] == (3, ih, iw), crops.shape
        hm = np.empty((P, 17, ih // 4, iw // 4), out_dtype)
        check(self._lib.hbp_hrnet_forward(self._ctx, ptr(crops), P, ptr(hm), _NP2HBP[np.dtype(out_dtype)], HOST))
        return hm

    # parity hooks (per-stage error attribution against the fp32 oracle)
    def hrnet_op_names(self):
        n = C.c_int()
        check(self._lib.hbp_hrnet_op_name(self._ctx, 0, None, 0, C.byref(n)))
        buf = C.create_string_buffer(256)
        names = []
        for i in range(n.value):
            check(self._lib.hbp_hrnet_op_name(self._ctx, i, buf, 256, None))
            names.append(buf.value.decode())
        return names

    def hrnet_forward_until(self, crops, op_index):
        crops = _c(crops, np.float16)
        check(self._lib.hbp_hrnet_forward_until(self._ctx, ptr(crops), crops.shape[0], int(op_index), HOST))

    def hrnet_debug_tensor(self, op_index):
        """output tensor of program op `op_index` after the last forward: (P,h,w,c) fp16 NHWC (c as stored)"""
        n, h, w, c = C.c_int(), C.c_int(), C.c_int(), C.c_int()
        check(self._lib.hbp_hrnet_debug_tensor(self._ctx, int(op_index), None, 0, C.byref(n), C.byref(h), C.byref(w), C.byref(c)))
        out = np.empty((n.value, h.value, w.value, c.value), np.float16)
        check(self._lib.hbp_hrnet_debug_tensor(self._ctx, int(op_index), ptr(out), out.nbytes, None, None, None, None))
        return out

    # ---- K6 ----------------------------------------------------------------
    def decode_proportions(self, heatmaps, boxes_yxyx_px=None, height_cm=None,
                           joint_thr=KEYPOINT_THRES_LIST, quarter_offset=False, mats=None, crop_hw=None):
        """mats (P,2,3) + crop_hw=(crop_h, crop_w): map the keypoints back through the crops' own dst->src matrices
        (general inverse affine) instead of the reference's box formula."""
        hm = np.ascontiguousarray(heatmaps)
        if hm.dtype not in (np.float32, np.float16):
            hm = hm.astype(np.float32)
        P, J, Hh, Wh = hm.shape
        out = dict(kpts_hm=np.zeros((P, J, 2), np.float32), scores=np.zeros((P, J), np.float32),
                   argmax=np.zeros((P, J), np.int32))
        boxes = hcm = thr = None
        if boxes_yxyx_px is not None:
            boxes = _c(boxes_yxyx_px, np.float32).reshape(P, 4)
            hcm = _c(np.broadcast_to(np.asarray(height_cm, np.float64), (P,)), np.float64)
            thr = _c(joint_thr, np.float32)
            out.update(kpts_img=np.zeros((P, J, 2), np.float32), ignored=np.zeros((P,), np.uint32))
            if J == 17:
                out.update(lengths_cm=np.zeros((P, 11), np.float32), torso_cm=np.zeros((P,), np.float64))
        if mats is not None:
            M = _c(mats, np.float64).reshape(P, 6)
            check(self._lib.hbp_decode_proportions_affine(
                self._ctx, ptr(hm), _NP2HBP[hm.dtype], P, J, Hh, Wh, ptr(boxes), ptr(M), int(crop_hw[0]), int(crop_hw[1]),
                ptr(hcm), ptr(thr), int(quarter_offset), ptr(out["kpts_hm"]), ptr(out.get("kpts_img")), ptr(out["scores"]),
                ptr(out["argmax"]), ptr(out.get("ignored")), ptr(out.get("lengths_cm")), ptr(out.get("torso_cm")), HOST))
            return out
        check(self._lib.hbp_decode_proportions(
            self._ctx, ptr(hm), _NP2HBP[hm.dtype], P, J, Hh, Wh, ptr(boxes), ptr(hcm), ptr(thr),
            int(quarter_offset), ptr(out["kpts_hm"]), ptr(out.get("kpts_img")), ptr(out["scores"]),
            ptr(out["argmax"]), ptr(out.get("ignored")), ptr(out.get("lengths_cm")), ptr(out.get("torso_cm")),
            HOST))
        return out

    def keypoint_lengths(self, kpts_img, pixel_to_cm, ignored=None):
        """Segment lengths from keypoints the caller already holds (reference modules/pose_estimator.py:130-200):
        kpts_img (P,17,2), pixel_to_cm scalar or (P,), ignored (P,) uint32 bit masks -> lengths_cm (P,11), torso_cm (P,)"""
        k = _c(kpts_img, np.float32).reshape(-1, 17, 2)
        P = k.shape[0]
        p2c = _c(np.broadcast_to(np.asarray(pixel_to_cm, np.float64), (P,)), np.float64)
        ign = None if ignored is None else _c(ignored, np.uint32).reshape(P)
        out = dict(lengths_cm=np.zeros((P, 11), np.float32), torso_cm=np.zeros((P,), np.float64))
        check(self._lib.hbp_keypoint_lengths(self._ctx, ptr(k), ptr(ign), ptr(p2c), P, ptr(out["lengths_cm"]),
                                             ptr(out["torso_cm"]), HOST))
        return out

    # ---- fused pipeline --------------------------------------------------------
    def pose_pipeline(self, frames, mats, frame_idx, boxes_yxyx_px, height_cm,
                      joint_thr=KEYPOINT_THRES_LIST, swap_rb=True, quarter_offset=False,
                      return_heatmaps=None):
        """frames (n,h,w,3) u8 + per-person 2x3 matrices/boxes -> keypoints, scores, lengths.
        return_heatmaps: None | np.float16 | np.float32."""
        frames = _c(frames, np.uint8)
        if frames.ndim == 3:
            frames = frames[None]
        mats = _c(mats, np.float64).reshape(-1, 6)
        P = mats.shape[0]
        fi = _c(frame_idx, np.int32)
        boxes = _c(boxes_yxyx_px, np.float32).reshape(P, 4)
        hcm = _c(np.broadcast_to(np.asarray(height_cm, np.float64), (P,)), np.float64)
        thr = _c(joint_thr, np.float32)
        _, ih, iw = self.hrnet
        out = dict(kpts_img=np.zeros((P, 17, 2), np.float32), scores=np.zeros((P, 17), np.float32),
                   ignored=np.zeros((P,), np.uint32), lengths_cm=np.zeros((P, 11), np.float32),
                   torso_cm=np.zeros((P,), np.float64))
        hm = None
        prm = _capi.PipelineParams(frames.shape[0], frames.shape[1], frames.shape[2], P, int(swap_rb),
                                   int(quarter_offset), F32)
        if return_heatmaps is not None:
            hm = np.empty((P, 17, ih // 4, iw // 4), return_heatmaps)
            prm.heatmap_dtype = _NP2HBP[np.dtype(return_heatmaps)]
        check(self._lib.hbp_pose_pipeline(self._ctx, C.byref(prm), ptr(frames), ptr(mats), ptr(fi), ptr(boxes),
                                          ptr(hcm), ptr(thr), ptr(out["kpts_img"]), ptr(out["scores"]),
                                          ptr(out["ignored"]), ptr(out["lengths_cm"]), ptr(out["torso_cm"]),
                                          ptr(hm)))
        if hm is not None:
            out["heatmaps"] = hm
        return out

    def pose_pipeline_submit(self, frames, mats, frame_idx, boxes_yxyx_px, height_cm,
                             joint_thr=KEYPOINT_THRES_LIST, swap_rb=True, quarter_offset=False):
        """Asynchronous form of pose_pipeline (hbp_pose_pipeline_submit): enqueue one batch and return a
        ticket; at most two tickets may be outstanding.  Submitting batch n+1 before collecting batch n overlaps
        its frame upload with the network of batch n.  `frames` should be pinned (Engine.pinned_empty) and must
        not be modified until the ticket is collected."""
        frames = _c(frames, np.uint8)
        if frames.ndim == 3:
            frames = frames[None]
        mats = _c(mats, np.float64).reshape(-1, 6)
        P = mats.shape[0]
        fi = _c(frame_idx, np.int32)
        boxes = _c(boxes_yxyx_px, np.float32).reshape(P, 4)
        hcm = _c(np.broadcast_to(np.asarray(height_cm, np.float64), (P,)), np.float64)
        thr = _c(joint_thr, np.float32)
        prm = _capi.PipelineParams(frames.shape[0], frames.shape[1], frames.shape[2], P, int(swap_rb),
                                   int(quarter_offset), F16)
        ticket = C.c_int(-1)
        check(self._lib.hbp_pose_pipeline_submit(self._ctx, C.byref(prm), ptr(frames), ptr(mats), ptr(fi), ptr(boxes),
                                                 ptr(hcm), ptr(thr), C.byref(ticket)))
        if not hasattr(self, "_inflight"):
            self._inflight = {}
        self._inflight[ticket.value] = (frames, P)          # keeps the frame buffer alive until collect
        return ticket.value

    def pose_pipeline_collect(self, ticket):
        """Wait for a submitted batch -> the dict pose_pipeline returns (without heatmaps)."""
        _, P = self._inflight.pop(ticket)
        out = dict(kpts_img=np.zeros((P, 17, 2), np.float32), scores=np.zeros((P, 17), np.float32),
                   ignored=np.zeros((P,), np.uint32), lengths_cm=np.zeros((P, 11), np.float32),
                   torso_cm=np.zeros((P,), np.float64))
        check(self._lib.hbp_pose_pipeline_collect(self._ctx, ticket, ptr(out["kpts_img"]), ptr(out["scores"]),
                                                  ptr(out["ignored"]), ptr(out["lengths_cm"]), ptr(out["torso_cm"])))
        return out


    # ---- chained det -> pose pipeline (configs[2] / [3]) -------------------------------------------------
    def _det_pose_params(self, frames, detector, persons_cap, swap_rb, quarter_offset, **kw):
        prm = _capi.DetPoseParams()
        prm.n_frames, prm.h, prm.w = frames.shape[0], frames.shape[1], frames.shape[2]
        prm.detector, prm.persons_cap = detector, int(persons_cap)
        prm.swap_rb, prm.quarter_offset = int(swap_rb), int(quarter_offset)
        for k, v in kw.items():
            setattr(prm, k, v)
        return prm

    def det_pose_submit_yolo(self, frames, pred, person_height=(175,), persons_cap=32, conf_thres=0.4, iou_thres=0.5,
                             person_class=0, in_size=(640, 640), resample="bilinear", max_det=300, cand_cap=4096,
                             joint_thr=KEYPOINT_THRES_LIST, swap_rb=False, quarter_offset=False):
        """frames (F,h,w,3) u8 RGB + decoded YOLOv5 head pred (F,N,5+nc) f32 -> ticket.  The whole chain letterbox ->
        NMS(person) -> scale_coords -> crop -> HRNet -> decode runs on the device (hbp_det_pose_submit)."""
        frames = _c(frames, np.uint8)
        if frames.ndim == 3:
            frames = frames[None]
        pred = _c(pred, np.float32)
        F_, N, E = pred.shape
        assert F_ == frames.shape[0]
        prm = self._det_pose_params(frames, _capi.DET_YOLO, persons_cap, swap_rb, quarter_offset, person_class=int(person_class),
                                    N=N, nc=E - 5, in_h=int(in_size[1]), in_w=int(in_size[0]),
                                    letterbox_mode=1 if resample == "bicubic" else 0, max_det=int(max_det),
                                    cand_cap=int(cand_cap), conf_thres=float(conf_thres), iou_thres=float(iou_thres))
        return self._det_pose_submit(prm, frames, pred, None, None, person_height, joint_thr)

    def det_pose_submit_edet(self, frames, boxes, scores, classes, person_height=(175,), persons_cap=None, det_threshold=0.70,
                             max_persons=3, x_expand=None, y_expand=0, person_class=1, joint_thr=KEYPOINT_THRES_LIST,
                             swap_rb=False, quarter_offset=False):
        """frames (F,h,w,3) u8 as the model sees them + EfficientDet outputs boxes (F,K,4) yxyx px, scores (F,K),
        classes (F,K) -> ticket.  x_expand defaults to h // 17 like the reference (person_det_pose_edet4_trtserver.py:116)."""
        frames = _c(frames, np.uint8)
        if frames.ndim == 3:
            frames = frames[None]
        boxes, scores, classes = _c(boxes, np.float32), _c(scores, np.float32), _c(classes, np.float32)
        if boxes.ndim == 2:
            boxes, scores, classes = boxes[None], scores[None], classes[None]
        F_, K = scores.shape
        assert F_ == frames.shape[0]
        if x_expand is None:
            x_expand = frames.shape[1] // 17
        if persons_cap is None:
            persons_cap = F_ * max_persons
        prm = self._det_pose_params(frames, _capi.DET_EDET, persons_cap, swap_rb, quarter_offset, person_class=int(person_class),
                                    K=K, max_persons=int(max_persons), det_thres=float(det_threshold),
                                    x_expand=float(x_expand), y_expand=float(y_expand))
        return self._det_pose_submit(prm, frames, boxes, scores, classes, person_height, joint_thr)

    def _det_pose_submit(self, prm, frames, d0, d1, d2, person_height, joint_thr):
        hts = _c(np.atleast_1d(np.asarray(person_height, np.float64)), np.float64)
        thr = _c(joint_thr, np.float32)
        ticket = C.c_int(-1)
        check(self._lib.hbp_det_pose_submit(self._ctx, C.byref(prm), ptr(frames), ptr(d0), ptr(d1), ptr(d2), ptr(hts), hts.size,
                                            ptr(thr), C.byref(ticket)))
        if not hasattr(self, "_inflight"):
            self._inflight = {}
        self._inflight[ticket.value] = ((frames, d0, d1, d2), int(prm.persons_cap))     # buffers stay alive until collect
        return ticket.value

    def det_pose_collect(self, ticket, return_heatmaps=False):
        """-> dict(n, status, frame_idx (n), boxes_yxyx_px (n,4), kpts_img (n,17,2), scores (n,17), ignored (n),
        lengths_cm (n,11), torso_cm (n) [, heatmaps (n,17,Hh,Wh) f16])"""
        _, cap = self._inflight.pop(ticket)
        out = dict(frame_idx=np.zeros(cap, np.int32), boxes_yxyx_px=np.zeros((cap, 4), np.float32),
                   kpts_img=np.zeros((cap, 17, 2), np.float32), scores=np.zeros((cap, 17), np.float32),
                   ignored=np.zeros(cap, np.uint32), lengths_cm=np.zeros((cap, 11), np.float32), torso_cm=np.zeros(cap, np.float64))
        n, status = C.c_int(0), C.c_int(0)
        hm = None
        if return_heatmaps:
            _, ih, iw = self.hrnet
            hm = np.zeros((cap, 17, ih // 4, iw // 4), np.float16)
        check(self._lib.hbp_det_pose_collect(self._ctx, ticket, C.byref(n), C.byref(status), ptr(out["frame_idx"]),
                                             ptr(out["boxes_yxyx_px"]), ptr(out["kpts_img"]), ptr(out["scores"]), ptr(out["ignored"]),
                                             ptr(out["lengths_cm"]), ptr(out["torso_cm"]), ptr(hm)))
        out = {k: v[:n.value] for k, v in out.items()}
        out["n"], out["status"] = n.value, status.value
        if hm is not None:
            out["heatmaps"] = hm[:n.value]
        return out


def lengths_to_dict(lengths_row, torso):
    """(11,) float32 + float64 torso -> the reference's dict
    (pose_estimator.py:191-200): np.float32 values, np.float64 torso, or the
    string "Part not visible" for 0."""
    d = {}
    for k, key in enumerate(SEGMENT_KEYS):
        v = lengths_row[k]
        if key == "torso":
            d[key] = np.float64(torso) if torso > 0 else NOT_VISIBLE
        else:
            d[key] = np.float32(v) if v > 0 else NOT_VISIBLE
    return d


_default = {}
_default_lock = threading.Lock()


def default_engine(device=0):
    with _default_lock:
        if device not in _default:
            _default[device] = Engine(device)
        return _default[device]


class MultiGpuEngine:
    """Frame-level data parallelism over the GPUs of one box:
    unit = frame (+ its persons); frame f -> GPU f mod G; results concatenated
    on the host in frame order.  No collective, no peer traffic."""

    def __init__(self, devices=None, engines=None, contexts_per_gpu=1, **hrnet_kw):
        """devices: GPU ids (default: all); contexts_per_gpu > 1 puts that many engine contexts on every listed GPU (the
        list may also name a GPU several times itself); engines: ready-made Engine objects instead."""
        if engines is not None:                     # ready-made engines (HRNet already loaded)
            self.engines = list(engines)
            self._hrnet_kw = hrnet_kw
            return
        if devices is None:
            n = C.c_int()
            check(_capi.lib().hbp_device_count(C.byref(n)))
            devices = list(range(n.value))
        devices = [d for d in devices for _ in range(max(1, int(contexts_per_gpu)))]
        self.engines = [Engine(d) for d in devices]
        self._hrnet_kw = hrnet_kw
        weights = None
        for e in self.engines:
            weights = e.load_hrnet(weights=weights, **hrnet_kw)

    @staticmethod
    def shard(n_frames, n_ranks):
        """frame indices per rank (round robin)"""
        return [list(range(r, n_frames, n_ranks)) for r in range(n_ranks)]

    def pose_pipeline(self, frames, mats, frame_idx, boxes_yxyx_px, height_cm, **kw):
        frames = np.asarray(frames)
        frame_idx = np.asarray(frame_idx, np.int32)
        mats = np.asarray(mats, np.float64).reshape(-1, 6)
        boxes = np.asarray(boxes_yxyx_px, np.float32).reshape(-1, 4)
        hcm = np.broadcast_to(np.asarray(height_cm, np.float64), (mats.shape[0],))
        G = len(self.engines)
        plan = self.shard(frames.shape[0], G)
        results = [None] * G
        errors = []

        def work(r):
            try:
                fr = plan[r]
                if not fr:
                    return
                remap = {f: i for i, f in enumerate(fr)}
                sel = np.nonzero(np.isin(frame_idx, fr))[0]
                if sel.size == 0:
                    return
                local_idx = np.array([remap[f] for f in frame_idx[sel]], np.int32)
                out = self.engines[r].pose_pipeline(frames[fr], mats[sel], local_idx, boxes[sel], hcm[sel], **kw)
                results[r] = (sel, out)
            except Exception as exc:      # surfaced after the join
                errors.append(exc)

        threads = [threading.Thread(target=work, args=(r,)) for r in range(G)]
        for t in threads:
            t.start()
        for t in threads:
            t.join()
        if errors:
            raise errors[0]
        merged = None
        P = mats.shape[0]
        for item in results:
            if item is None:
                continue
            sel, out = item
            if merged is None:
                merged = {k: np.zeros((P,) + v.shape[1:], v.dtype) for k, v in out.items()}
            for k, v in out.items():
                merged[k][sel] = v
        return merged

    def det_pose_stream(self, submit, n_steps, depth=2, collect_kw=None):
        """Chained det->pose steps (hbp_det_pose_submit/_collect) round robin over the engines, `depth` (<= 2) tickets in
        flight per engine, driven from the calling thread (submits only enqueue).  `submit(engine, rank, step)` -> ticket, e.g.
        `lambda e, r, s: e.det_pose_submit_yolo(frame[r][s & 1], head[r][s & 1], persons_cap=64)` with per-engine pinned
        buffers.  The engines may be several contexts on ONE GPU -- `MultiGpuEngine(devices=[0, 0])`: each context has its
        own streams, activation buffers and graphs, so the forward of one step overlaps the tail of the previous one
        (+8 % steps/s on a B200: tools/dual_engine_e2e.py).  Returns the results in step order."""
        G = len(self.engines)
        results = [None] * n_steps
        inflight = []                               # (step, engine, ticket)
        kw = collect_kw or {}
        for s in range(n_steps):
            e = self.engines[s % G]
            inflight.append((s, e, submit(e, s % G, s // G)))
            if len(inflight) >= depth * G:
                g, eg, tk = inflight.pop(0)
                results[g] = eg.det_pose_collect(tk, **kw)
        for g, eg, tk in inflight:
            results[g] = eg.det_pose_collect(tk, **kw)
        return results

    def stream(self, frame_source, n_frames, depth=2):
        """Frame stream sharded over the GPUs (BASELINE configs[4]): frame f goes to GPU f mod G, every GPU keeps
        `depth` (<= 2) frames in flight through pose_pipeline_submit/_collect, one host thread per GPU.
        `frame_source(f)` -> (frame_u8 (h,w,3) [ideally pinned], mats (P,6), boxes_yxyx_px (P,4), height_cm).
        Returns (results in frame order, per-frame latency in ms measured submit -> collect)."""
        import time
        G = len(self.engines)
        results, latency = [None] * n_frames, [0.0] * n_frames
        errors = []

        def work(r):
            try:
                eng = self.engines[r]
                pending = []                        # (frame index, ticket, t_submit)
                for f in range(r, n_frames, G):
                    frame, mats, boxes, hcm = frame_source(f)
                    fi = np.zeros(len(boxes), np.int32)
                    t0 = time.perf_counter()
                    pending.append((f, eng.pose_pipeline_submit(frame, mats, fi, boxes, hcm), t0))
                    if len(pending) >= depth:
                        g, tk, ts = pending.pop(0)
                        results[g] = eng.pose_pipeline_collect(tk)
                        latency[g] = (time.perf_counter() - ts) * 1e3
                for g, tk, ts in pending:
                    results[g] = eng.pose_pipeline_collect(tk)
                    latency[g] = (time.perf_counter() - ts) * 1e3
            except Exception as exc:      # surfaced after the join
                errors.append(exc)

        threads = [threading.Thread(target=work, args=(r,)) for r in range(G)]
        for t in threads:
            t.start()
        for t in threads:
            t.join()
        if errors:
            raise errors[0]
        return results, latency
