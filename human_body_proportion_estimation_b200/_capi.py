"""ctypes binding of libhbp_b200.so (include/hbp.h).

The library is the product: when it is missing this module raises, it never
falls back to numpy.  ctypes releases the GIL for the duration of every call,
so N Python threads can drive N contexts (one per GPU).
"""
import ctypes as C
import os

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(HERE, "libhbp_b200.so")

HOST, DEVICE = 0, 1
U8, F16, F32 = 0, 1, 2
NCHW, NHWC = 0, 1
PRE_COPY, PRE_STRETCH, PRE_LETTERBOX, PRE_LETTERBOX_PIL = 0, 1, 2, 3

_lib = None


class HbpError(RuntimeError):
    pass


class PipelineParams(C.Structure):
    _fields_ = [("n_frames", C.c_int), ("h", C.c_int), ("w", C.c_int), ("P", C.c_int),
                ("swap_rb", C.c_int), ("quarter_offset", C.c_int), ("heatmap_dtype", C.c_int)]


class DetPoseParams(C.Structure):
    _fields_ = [("n_frames", C.c_int), ("h", C.c_int), ("w", C.c_int), ("detector", C.c_int), ("persons_cap", C.c_int),
                ("swap_rb", C.c_int), ("quarter_offset", C.c_int), ("person_class", C.c_int),
                ("N", C.c_int), ("nc", C.c_int), ("in_h", C.c_int), ("in_w", C.c_int), ("letterbox_mode", C.c_int),
                ("max_det", C.c_int), ("cand_cap", C.c_int), ("conf_thres", C.c_float), ("iou_thres", C.c_double),
                ("K", C.c_int), ("max_persons", C.c_int), ("det_thres", C.c_float), ("x_expand", C.c_float),
                ("y_expand", C.c_float), ("mem", C.c_int)]


DET_YOLO, DET_EDET = 0, 1

_P = C.c_void_p
_I = C.c_int
_SIGS = {
    "hbp_version": (C.c_int, []),
    "hbp_last_error": (C.c_char_p, []),
    "hbp_device_count": (_I, [C.POINTER(_I)]),
    "hbp_ctx_create": (_I, [_I, C.POINTER(_P)]),
    "hbp_ctx_destroy": (_I, [_P]),
    "hbp_sync": (_I, [_P]),
    "hbp_dev_alloc": (_I, [_P, C.c_size_t, C.POINTER(_P)]),
    "hbp_dev_free": (_I, [_P, _P]),
    "hbp_host_alloc": (_I, [_P, C.c_size_t, C.POINTER(_P)]),
    "hbp_host_free": (_I, [_P, _P]),
    "hbp_copy_h2d": (_I, [_P, _P, _P, C.c_size_t]),
    "hbp_copy_d2h": (_I, [_P, _P, _P, C.c_size_t]),
    "hbp_memset_dev": (_I, [_P, _P, _I, C.c_size_t]),
    "hbp_timer_start": (_I, [_P, _I]),
    "hbp_timer_stop": (_I, [_P, _I]),
    "hbp_timer_elapsed_ms": (_I, [_P, _I, C.POINTER(C.c_float)]),
    "hbp_flush_l2": (_I, [_P]),
    "hbp_kernel_launches": (_I, [_P, C.POINTER(C.c_uint64)]),
    "hbp_preprocess": (_I, [_P, _P, _I, _I, _I, _I, _I, _I, _I, _I, _P, _I, _I, _I]),
    "hbp_yolo_decode_raw": (_I, [_P, _P, _P, _P, _I, _I, _I, _I, _I, _I, _I, _P, _I]),
    "hbp_yolo_nms": (_I, [_P, _P, _I, _I, _I, C.c_float, C.c_double, _P, _I, _I, _P, _P, _I]),
    "hbp_yolo_filter": (_I, [_P, _P, _I, _I, _I, C.c_float, _P, _I, _I, _P, _I]),
    "hbp_yolo_nms_legacy": (_I, [_P, _P, _I, _I, _I, C.c_float, C.c_float, _I, _P, _P, _I]),
    "hbp_scale_coords": (_I, [_P, _P, _I, _I, _I, _I, _I, _I]),
    "hbp_edet_person_filter": (_I, [_P, _P, _P, _P, _I, _I, C.c_float, C.c_float, C.c_float,
                                    C.c_float, _I, _I, _I, _P, _P, _I]),
    "hbp_crop_warp": (_I, [_P, _P, _I, _I, _I, _P, _P, _I, _I, _I, _I, _P, _I, _I]),
    "hbp_hrnet_describe": (_I, [_I, _I, _I, C.c_char_p, C.c_size_t, C.POINTER(C.c_size_t),
                                C.POINTER(C.c_size_t), C.POINTER(C.c_size_t)]),
    "hbp_hrnet_load": (_I, [_P, _I, _I, _I, _P, C.c_size_t, _P, C.c_size_t]),
    "hbp_hrnet_forward": (_I, [_P, _P, _I, _P, _I, _I]),
    "hbp_conv2d_nhwc": (_I, [_P, _I, _P, _I, _I, _I, _I, _P, _P, _P, _I, _I, _I, _I, _I, _P, C.POINTER(_I), _I]),
    "hbp_conv2d_nhwc_timed": (_I, [_P, _I, _P, _I, _I, _I, _I, _P, _P, _P, _I, _I, _I, _I, _I, _P, C.POINTER(_I), _I, _I,
                                   C.POINTER(C.c_float)]),
    "hbp_hrnet_set_engine": (_I, [_P, _I]),
    "hbp_hrnet_debug_tensor": (_I, [_P, _I, _P, C.c_size_t, C.POINTER(_I), C.POINTER(_I),
                                    C.POINTER(_I), C.POINTER(_I)]),
    "hbp_hrnet_forward_until": (_I, [_P, _P, _I, _I, _I]),
    "hbp_hrnet_op_name": (_I, [_P, _I, C.c_char_p, C.c_size_t, C.POINTER(_I)]),
    "hbp_decode_proportions": (_I, [_P, _P, _I, _I, _I, _I, _I, _P, _P, _P, _I, _P, _P, _P, _P, _P,
                                    _P, _P, _I]),
    "hbp_decode_proportions_affine": (_I, [_P, _P, _I, _I, _I, _I, _I, _P, _P, _I, _I, _P, _P, _I, _P, _P, _P, _P, _P,
                                           _P, _P, _I]),
    "hbp_keypoint_lengths": (_I, [_P, _P, _P, _P, _I, _P, _P, _I]),
    "hbp_det_pose_submit": (_I, [_P, C.POINTER(DetPoseParams), _P, _P, _P, _P, _P, _I, _P, C.POINTER(_I)]),
    "hbp_det_pose_collect": (_I, [_P, _I, C.POINTER(_I), C.POINTER(_I), _P, _P, _P, _P, _P, _P, _P, _P]),
    "hbp_pose_pipeline": (_I, [_P, C.POINTER(PipelineParams), _P, _P, _P, _P, _P, _P, _P, _P, _P, _P,
                               _P, _P]),
    "hbp_pose_pipeline_submit": (_I, [_P, C.POINTER(PipelineParams), _P, _P, _P, _P, _P, _P, C.POINTER(_I)]),
    "hbp_pose_pipeline_collect": (_I, [_P, _I, _P, _P, _P, _P, _P]),
}
EXPORTS = tuple(_SIGS)


def lib():
    """Load (once) and return the shared library; raise if it is not built."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise HbpError(
                "CUDA library %s is missing: build it with "
                "`python -m human_body_proportion_estimation_b200.build` "
                "(there is no CPU fallback)" % LIB_PATH)
        l = C.CDLL(LIB_PATH)
        for name, (res, args) in _SIGS.items():
            fn = getattr(l, name)
            fn.restype = res
            fn.argtypes = args
        _lib = l
    return _lib


def check(status):
    if status != 0:
        raise HbpError("hbp error %d: %s" % (status, lib().hbp_last_error().decode("utf-8", "replace")))


def ptr(a):
    """numpy array (C-contiguous) -> void*; None -> NULL; ints pass through (device pointers)."""
    if a is None:
        return None
    if isinstance(a, (int, np.integer)):
        return C.c_void_p(int(a))
    if not a.flags["C_CONTIGUOUS"]:
        raise ValueError("array must be C-contiguous")
    return C.c_void_p(a.ctypes.data)


def describe_hrnet(width, in_h, in_w, full=False):
    """[(name, cin, cout, k, stride, w_off, b_off)], n_weights, n_biases -- host only.
    full=True appends (out_h, out_w, up) to every row."""
    l = lib()
    nw, nb, need = C.c_size_t(), C.c_size_t(), C.c_size_t()
    st = l.hbp_hrnet_describe(width, in_h, in_w, None, 0, C.byref(nw), C.byref(nb), C.byref(need))
    check(st)
    buf = C.create_string_buffer(need.value)
    check(l.hbp_hrnet_describe(width, in_h, in_w, buf, need.value, None, None, None))
    rows = []
    for line in buf.value.decode().splitlines():
        f = line.split()
        row = (f[0], int(f[1]), int(f[2]), int(f[3]), int(f[4]), int(f[5]), int(f[6]))
        rows.append(row + (int(f[7]), int(f[8]), int(f[9])) if full else row)
    return rows, nw.value, nb.value
