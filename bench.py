#!/usr/bin/env python
"""Benchmark of the top-down pose hot path (BASELINE.json metric: person crops/sec
through det -> crop -> HRNet -> decode).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--config 1|2|3]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...

One step = ONE call of the chained pipeline (hbp_det_pose_submit / _collect): frames + detector-head tensors ->
letterbox (K1) -> person filter + NMS (K2/K3) -> scale_coords + per-person crop parameters -> crop (K4) ->
HRNet (K5, one CUDA graph) -> decode + proportions (K6), no host round trip between the stages.  The detector
BACKBONES are not part of the reference tree (Google-Drive artifacts), so their output tensors are synthetic inputs.

Workloads (config.workload):
  --config 1 (default, the headline)  BASELINE configs[1]: HRNet-W32 256x192 fp16, 64 person crops from one 1080p frame
              per step and GPU; the 64 persons come out of the YOLO head stage (synthetic decoded head whose NMS keeps 64)
  --config 2  BASELINE configs[2]: YOLOv5s 640x640 head (30 planted persons + distractors) -> HRNet-W32, per frame
  --config 3  BASELINE configs[3]: EfficientDet outputs -> HRNet-W48 384x288, 16 frames x 16 persons per step

  value     crops/s with frames and head tensors resident in HBM (params.mem = HBP_DEVICE), device-timed with CUDA
            events on the library's stream, L2 flushed between steps
  e2e       crops/s through the public API (Engine.det_pose_submit_* / det_pose_collect) with HOST buffers: pinned frame
            and head tensors -> H2D -> chain -> D2H of the results, two batches in flight
  roofline  HRNet conv stack: algorithmic FLOPs (2 x MACs of the 293 convs x person slots computed) / CUDA-event time
            of the HRNet stage, against the measured sustained bf16 peak of MEASURED_PEAKS.json
  cpu_baseline / --impl reference
            the reference's CPU path on the box's host cores, from oracle/ only (never loads libhbp_b200.so): PIL
            letterbox, torchvision NMS via the reference's algorithm restated, cv2.warpAffine crops, torch fp32 HRNet
            (stand-in for onnxruntime, which is not installable offline), the reference's per-person decode loop
After the timed regions the e2e results are checked against the oracle on the device's own heatmaps.
"""
import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

FRAME_H, FRAME_W = 1080, 1920
METRIC = "person crops/sec (det->crop->HRNet->decode)"
UNIT = "crops/s"
HEIGHTS = [175.0]

CONFIGS = {
    1: dict(name="configs[1]", detector="yolo", width=32, in_h=256, in_w=192, frames=1, persons=64, cap=64,
            text="configs[1]: HRNet-W32 256x192 fp16, 64 synthetic person crops from one 1080p frame per step per GPU "
                 "(letterbox -> YOLO-head NMS(person) -> scale_coords -> crop -> HRNet -> decode+proportions)"),
    2: dict(name="configs[2]", detector="yolo", width=32, in_h=256, in_w=192, frames=1, persons=None, cap=48,
            text="configs[2]: YOLOv5s 640x640 head (25200x85, 30 planted persons + 300 distractors) + NMS feeding HRNet-W32 "
                 "256x192 crops, end to end per 1080p frame (48 person slots computed)"),
    3: dict(name="configs[3]", detector="edet", width=48, in_h=384, in_w=288, frames=16, persons=256, cap=256,
            text="configs[3]: EfficientDet outputs (100 rows/frame) -> person filter -> HRNet-W48 384x288 -> proportions, "
                 "batch of 16 1080p frames x 16 persons"),
}


def synth_inputs(cfg):
    """(frames (F,h,w,3) u8 RGB, detector tensors tuple) -- numpy only, shared by both arms"""
    from human_body_proportion_estimation_b200 import synth
    F = cfg["frames"]
    frames = np.stack([synth.frame_u8(FRAME_H, FRAME_W, seed=synth.SEED_BASE + 2 + i) for i in range(F)])
    if cfg["detector"] == "yolo":
        pred = synth.yolo_head_grid()[0] if cfg["persons"] == 64 else synth.yolo_decoded_head()[0]
        return frames, (pred,)
    b, s, c = synth.edet_outputs(F, 16, FRAME_H, FRAME_W)
    return frames, (b, s, c)


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return dict(tflops=d.get("bf16_tflops_sustained", d.get("bf16_tflops")), hbm=d.get("hbm_gbs"),
                    src="measured (MEASURED_PEAKS.json, sustained bf16)")
    return dict(tflops=1400.0, hbm=6650.0, src="fallback (B200_PROFILING.md)")


class ClockSampler:
    """nvidia-smi clocks + throttle reasons during the timed region"""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index, period_ms=100):
        self.proc, self.lines = None, []
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", "-i", str(index), "--query-gpu=" + self.Q, "--format=csv,noheader,nounits",
                 "-lms", str(period_ms)], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
            t0 = time.time()
            while not self.lines and time.time() - t0 < 3.0:       # first sample before the timed region starts
                time.sleep(0.01)
        except Exception:
            self.proc = None
        self.first = len(self.lines)

    def mark(self):
        self.first = len(self.lines)                   # samples from here on belong to the timed region

    def _read(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        names = ("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap")
        region = self.lines[self.first:] or self.lines[-1:]
        for ln in region:
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 7:
                continue
            try:
                sm.append(float(f[0])); mx.append(float(f[1]))
            except ValueError:
                continue
            for nm, v in zip(names, f[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(nm)
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "samples": len(sm), "reasons": sorted(reasons)}


# --------------------------------------------------------------------------
# CPU reference arm / cpu_baseline: oracle/ only, never the library under test
# --------------------------------------------------------------------------
def cpu_reference_run(cfg, steps, warmup, frames=None, dets=None):
    """The reference's CPU path on the config's workload, one full batch per step.
    Returns (crops_per_s, ms_per_step, cores, note)."""
    import torch
    from oracle import detect as od
    from oracle import geometry as og
    from oracle import hrnet_table, imgproc
    from oracle.hrnet_fp32 import HRNetFP32
    import cv2
    if frames is None:
        frames, dets = synth_inputs(cfg)
    if cfg["width"] == 48 and frames.shape[0] > 2:       # bounded sample: 2 of the 16 frames (32 of the 256 W48 crops) per step
        frames, dets = frames[:2], tuple(d[:2] for d in dets)
    net = HRNetFP32(hrnet_table.random_weights(cfg["width"], seed=0), cfg["width"])
    try:
        avail = len(os.sched_getaffinity(0))
    except Exception:
        avail = os.cpu_count() or 1
    if torch.get_num_threads() < avail:      # torchrun exports OMP_NUM_THREADS=1 for every rank
        torch.set_num_threads(avail)
    cores = torch.get_num_threads()
    cv2.setNumThreads(cores)
    ih, iw = cfg["in_h"], cfg["in_w"]
    H, W = frames.shape[1:3]

    def step():
        crops, boxes_px, hts = [], [], []
        for f in range(frames.shape[0]):
            if cfg["detector"] == "yolo":
                imgproc.letterbox_pil(frames[f], 640, 640)                       # obj_det_yolov5_onnx.py:27-36 (PIL bicubic)
                det = od.official_nms(dets[0][f:f + 1], 0.4, 0.5, classes=[0])[0]
                b = od.scale_coords((640, 640), det[:, :4].copy(), (H, W))
                for i, bb in enumerate(b):
                    x1, y1, x2, y2 = (int(v) for v in bb)
                    M = imgproc.box_resize_matrix((x1, y1, x2, y2), ih, iw)
                    c = cv2.warpAffine(frames[f], M, (iw, ih), flags=cv2.INTER_LINEAR | cv2.WARP_INVERSE_MAP,
                                       borderMode=cv2.BORDER_CONSTANT, borderValue=0)
                    crops.append(np.transpose(c / 255.0, (2, 0, 1)).astype(np.float32))
                    boxes_px.append(np.array([y1, x1, y2, x2], np.float32))
                    hts.append(HEIGHTS[min(i, len(HEIGHTS) - 1)])
            else:
                bn, _ = od.edet_person_filter(dets[0][f], dets[1][f], dets[2][f], 0.70, H // 17, 0, H, W, 16)
                for i, b in enumerate(bn):
                    M = imgproc.crop_and_resize_matrix(b, H, W, ih, iw)
                    c = cv2.warpAffine(frames[f], M, (iw, ih), flags=cv2.INTER_LINEAR | cv2.WARP_INVERSE_MAP,
                                       borderMode=cv2.BORDER_CONSTANT, borderValue=0)
                    crops.append(np.transpose(c / 255.0, (2, 0, 1)).astype(np.float32))
                    boxes_px.append(b * np.array([H, W, H, W], np.float32))
                    hts.append(HEIGHTS[min(i, len(HEIGHTS) - 1)])
        hm = net(np.stack(crops)).numpy()                       # torch fp32, all host threads
        out = [og.person_postprocess(hm[p], boxes_px[p], hts[p])["lengths"] for p in range(len(crops))]
        return len(out)

    for _ in range(warmup):
        step()
    t0 = time.perf_counter()
    n = 0
    for _ in range(steps):
        n += step()
    dt = time.perf_counter() - t0
    note = ("one full batch of the workload per step (%d crops): PIL-bicubic letterbox / NMS or EfficientDet person filter "
            "(oracle restatement of the reference), cv2.warpAffine crops, torch-fp32 HRNet-W%d on %d threads (stand-in for "
            "onnxruntime CPU), the reference's per-person decode + proportions loop" % (n // max(steps, 1), cfg["width"], cores))
    return n / dt, dt / steps * 1e3, cores, note


def run_reference(args, rank):
    if rank != 0:
        return
    cfg = CONFIGS[args.config]
    # bounded: the CPU path needs ~1 s per 64-crop W32 batch; cap the number of timed steps so the run ends in minutes
    steps = max(1, min(args.steps, 10 if cfg["width"] == 32 else 2))
    warm = max(1, min(args.warmup, 2 if cfg["width"] == 32 else 1))
    val, ms, cores, note = cpu_reference_run(cfg, steps, warm)
    line = {
        "impl": "reference", "metric": METRIC, "value": val, "unit": UNIT, "n_gpus": args.gpus,
        "steps": steps, "warmup": warm, "ms_per_step": ms, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": config_dict(cfg),
        "cpu_baseline": {"value": val, "unit": UNIT, "cores": cores, "kind": "port", "sample": note + "; %d timed steps" % steps},
        "e2e": {"value": val, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    emit(json.dumps(line))


def hrnet_traffic():
    """DRAM bytes per HRNet launch from the committed ncu launch list (profiles/*_hrnet_traffic.json, written by
    tools/launch_summary.py): bench.py cannot run ncu on itself."""
    best = None
    try:
        for f in sorted(os.listdir(os.path.join(ROOT, "profiles"))):
            if f.endswith("_hrnet_traffic.json"):
                best = os.path.join(ROOT, "profiles", f)
        if best:
            d = json.load(open(best))
            d["note"] = "dram__bytes_read+write per launch, average over the %d launches of one forward (%s)" % (
                round(d["launches_per_forward"]), os.path.basename(best))
            return d
    except Exception:
        pass
    return {}


def config_dict(cfg):
    return {"workload": cfg["text"], "frame": [FRAME_H, FRAME_W, 3], "frames_per_step_per_gpu": cfg["frames"],
            "person_slots_per_step_per_gpu": cfg["cap"],
            "hrnet": "W%d %dx%d" % (cfg["width"], cfg["in_h"], cfg["in_w"]), "weights": "random-init (seed 0), BN folded",
            "detector_outputs": "synthetic (the backbones are opaque artifacts outside the reference tree)",
            "l2": "value / e2e: the steps cycle through resident input sets (frame + detector head) that together exceed the 126 MB L2; single_context, roofline and stage legs: L2 flushed before every timed step (256 MiB write)",
            "parallelism": "frame-sharded, no collective"}


# --------------------------------------------------------------------------
# memory-bound stages on batched launches (at the config shapes K1/K2/K4/K6 move a
# few MB per call = microseconds of HBM time, below launch latency; their roofline fraction is
# measured on launches sized to ~1 GB, algorithmic bytes / CUDA-event time / measured HBM peak)
# --------------------------------------------------------------------------
def stage_rooflines(eng, hbm_gbs):
    import ctypes as C
    from human_body_proportion_estimation_b200 import geometry, synth
    from human_body_proportion_estimation_b200._capi import DEVICE, F16, NCHW, NHWC, PRE_COPY, PRE_LETTERBOX, PRE_LETTERBOX_PIL, U8, check
    from human_body_proportion_estimation_b200.engine import KEYPOINT_THRES_LIST
    lib, ctx = eng._lib, eng._ctx
    out = {}

    def timed(fn, reps=5):
        fn(); fn(); eng.sync()
        ms = []
        for _ in range(reps):
            eng.flush_l2()
            eng.timer_start(3)
            fn()
            eng.timer_stop(3)
            ms.append(eng.timer_ms(3))
        return min(ms)

    def entry(name, nbytes, ms, note):
        gbs = nbytes / (ms * 1e-3) / 1e9
        out[name] = {"bound": "hbm", "algorithmic_bytes": int(nbytes), "ms": ms, "achieved": gbs, "peak": hbm_gbs,
                     "unit": "GB/s", "frac": gbs / hbm_gbs, "launch": note}

    # K1: 64 4K frames u8 NHWC -> BGR->RGB copy (A1) and letterbox 640x640 f16 NCHW (A2)
    nf, fh, fw = 64, 2160, 3840
    frame = synth.frame_u8(fh, fw, seed=synth.SEED_BASE + 5, smooth=False)
    d_frames = eng.dev_alloc(nf * frame.nbytes)
    for i in range(nf):
        eng.h2d(d_frames + i * frame.nbytes, frame)
    d_copy = eng.dev_alloc(nf * frame.nbytes)
    d_lb = eng.dev_alloc(nf * 3 * 640 * 640 * 2)
    ms = timed(lambda: check(lib.hbp_preprocess(ctx, C.c_void_p(d_frames), nf, fh, fw, PRE_COPY, fh, fw, 1, 128,
                                                C.c_void_p(d_copy), U8, NHWC, DEVICE)))
    entry("k1_bgr2rgb_copy_64x4k_u8", 2 * nf * frame.nbytes, ms, "64 frames 2160x3840x3 u8, one launch")
    # the 6:1 bilinear downscale taps 2 of every 6 source rows; inside a tapped row the taps are 18 bytes apart, below the
    # 32-byte DRAM sector, so the whole row is fetched: bytes = tapped rows x row bytes + output
    nh = int(fh * min(640 / fw, 640 / fh))
    tapped = 2 * nh * fw * 3
    ms = timed(lambda: check(lib.hbp_preprocess(ctx, C.c_void_p(d_frames), nf, fh, fw, PRE_LETTERBOX, 640, 640, 1, 128,
                                                C.c_void_p(d_lb), F16, NCHW, DEVICE)))
    entry("k1_letterbox_64x4k_to_640_f16", nf * (tapped + 3 * 640 * 640 * 2), ms,
          "64 frames 2160x3840x3 u8 -> 3x640x640 f16, one launch (bytes = the %d source rows the bilinear taps touch, whole rows, + output)" % (2 * nh))
    ms = timed(lambda: check(lib.hbp_preprocess(ctx, C.c_void_p(d_frames), nf, fh, fw, PRE_LETTERBOX_PIL, 640, 640, 1, 128,
                                                C.c_void_p(d_lb), F16, NCHW, DEVICE)))
    entry("k1_letterbox_pil_bicubic_64x4k_to_640_f16", nf * (frame.nbytes + 3 * 640 * 640 * 2), ms,
          "64 frames 2160x3840x3 u8 -> 3x640x640 f16 with PIL's antialiased bicubic (every source pixel is a tap: bytes = whole frame + output)")
    eng.dev_free(d_copy); eng.dev_free(d_lb)

    # K4: 4096 crops (64 per frame) from the 64 4K frames -> (4096,3,256,192) f16
    P = 4096
    boxes = synth.person_boxes_yxyx_px(P, fh, fw, seed=synth.SEED_BASE + 6, hmin=300, hmax=1400)
    mats = geometry.crop_and_resize_matrices(boxes / np.array([fh, fw, fh, fw], np.float32), fh, fw, 256, 192)
    fidx = (np.arange(P) // 64).astype(np.int32)
    d_m = eng.to_device(mats.reshape(P, 6)); d_fi = eng.to_device(fidx)
    d_cr = eng.dev_alloc(P * 3 * 256 * 192 * 2)
    ms = timed(lambda: check(lib.hbp_crop_warp(ctx, C.c_void_p(d_frames), nf, fh, fw, C.c_void_p(d_m), C.c_void_p(d_fi),
                                               P, 256, 192, 1, C.c_void_p(d_cr), F16, DEVICE)))
    bw = np.clip(boxes[:, 3] - boxes[:, 1], 1, None); bh = np.clip(boxes[:, 2] - boxes[:, 0], 1, None)
    src = float(np.minimum(bw * bh, 4.0 * 256 * 192).sum()) * 3          # min(box area, 4 taps per output pixel) x 3 B
    entry("k4_crop_4096x256x192_f16", src + P * 3 * 256 * 192 * 2, ms, "4096 crops from 64 4K frames, one launch")
    eng.dev_free(d_frames); eng.dev_free(d_m); eng.dev_free(d_fi); eng.dev_free(d_cr)

    # K2: candidate filter on 128 decoded heads (25200 x 85 f32 = 8.568 MB each, 1.1 GB): hbp_yolo_nms with a threshold no
    # row passes, so that only the filter (+ the empty gather) runs
    B = 128
    pred = synth.yolo_decoded_head()[0]
    d_pred = eng.dev_alloc(B * pred.nbytes)
    for i in range(B):
        eng.h2d(d_pred + i * pred.nbytes, pred)
    d_det = eng.dev_alloc(B * 300 * 6 * 4); d_cnt = eng.dev_alloc(B * 4)
    d_cls = eng.to_device(np.zeros(1, np.int32))
    ms = timed(lambda: check(lib.hbp_yolo_filter(ctx, C.c_void_p(d_pred), B, 25200, 80, 0.4, C.c_void_p(d_cls), 1, 4096,
                                                 C.c_void_p(d_cnt), DEVICE)))
    cnt = np.zeros(B, np.int32)
    eng.d2h(cnt, d_cnt); eng.sync()
    # what the filter has to move: the objectness column (one 32-byte sector of every 340-byte row) and the full rows of the
    # candidates (obj > conf) -- the reference gathers exactly those rows (onnx_utils.py:133,168)
    entry("k2_yolo_filter_128x25200x85_f32", B * 25200 * 32 + int(cnt.sum()) * 85 * 4, ms,
          "128 decoded heads (1.1 GB resident), one filter launch, %d candidates per head: bytes = one sector per row + the candidates' rows" % int(cnt[0]))
    eng.dev_free(d_cls)
    eng.dev_free(d_pred); eng.dev_free(d_det); eng.dev_free(d_cnt)

    # K6: decode + proportions on 8192 crops of (17,64,48) f16
    P = 8192
    hm = synth.heatmaps(64, dtype=np.float16)
    d_hm = eng.dev_alloc(P * hm[0].nbytes)
    for i in range(P // 64):
        eng.h2d(d_hm + i * hm.nbytes, hm)
    bx = synth.person_boxes_yxyx_px(P, 1080, 1920, seed=synth.SEED_BASE + 7)
    d_bx = eng.to_device(bx); d_h = eng.to_device(np.full(P, 175.0)); d_t = eng.to_device(np.asarray(KEYPOINT_THRES_LIST, np.float32))
    d_kp = eng.dev_alloc(P * 17 * 8); d_sc = eng.dev_alloc(P * 17 * 4); d_ig = eng.dev_alloc(P * 4)
    d_ln = eng.dev_alloc(P * 44); d_to = eng.dev_alloc(P * 8)
    ms = timed(lambda: check(lib.hbp_decode_proportions(ctx, C.c_void_p(d_hm), F16, P, 17, 64, 48, C.c_void_p(d_bx), C.c_void_p(d_h),
                                                        C.c_void_p(d_t), 0, None, C.c_void_p(d_kp), C.c_void_p(d_sc), None,
                                                        C.c_void_p(d_ig), C.c_void_p(d_ln), C.c_void_p(d_to), DEVICE)))
    entry("k6_decode_proportions_8192x17x64x48_f16", P * (17 * 64 * 48 * 2 + 17 * 12 + 11 * 4 + 4), ms, "8192 crops, one launch")
    for d in (d_hm, d_bx, d_h, d_t, d_kp, d_sc, d_ig, d_ln, d_to):
        eng.dev_free(d)
    return out


def check_against_oracle(out, heights):
    """the chain's decode / remap / gate / lengths on ITS OWN heatmaps equal the oracle's per-person loop (checker only,
    outside every timed region)"""
    from oracle import geometry as og
    n = out["n"]
    for i in range(n):
        # person index inside its frame decides the height (person_det_pose_edet4_trtserver.py:166-168)
        j = int((out["frame_idx"][:i] == out["frame_idx"][i]).sum())
        r = og.person_postprocess(out["heatmaps"][i].astype(np.float32), out["boxes_yxyx_px"][i], heights[min(j, len(heights) - 1)])
        if not np.array_equal(out["kpts_img"][i], r["xy_img"]):
            return "keypoints of person %d differ from the oracle" % i
        if int(out["ignored"][i]) != sum(1 << k for k in r["ignored"]):
            return "ignored set of person %d differs from the oracle" % i
        got = out["lengths_cm"][i].astype(np.float64)
        got[1] = out["torso_cm"][i]
        if not np.array_equal(got, og.lengths_to_array(r["lengths"])):
            return "lengths of person %d differ from the oracle" % i
    return "ok (%d persons: keypoints, ignored sets and 11 lengths bit-exact vs oracle.geometry on the device's heatmaps)" % n


# --------------------------------------------------------------------------
# our arm
# --------------------------------------------------------------------------
def run_ours(args, rank, world, local_rank):
    import ctypes as C
    from human_body_proportion_estimation_b200 import _capi, hrnet_arch
    from human_body_proportion_estimation_b200.engine import Engine, KEYPOINT_THRES_LIST
    from human_body_proportion_estimation_b200._capi import DEVICE, F16, NCHW, PRE_LETTERBOX, PRE_LETTERBOX_PIL, check, ptr
    from human_body_proportion_estimation_b200 import dist_util
    cfg = CONFIGS[args.config]
    if world > 1:
        import torch
        torch.cuda.set_device(local_rank)
    grp = dist_util.Group(backend="nccl", device="cuda" if world > 1 else None)
    barrier = grp.barrier

    eng = Engine(local_rank)
    lib, ctx = eng._lib, eng._ctx
    hr_weights = eng.load_hrnet(None, cfg["width"], cfg["in_h"], cfg["in_w"], seed=0)
    frames, dets = synth_inputs(cfg)
    cap, F = cfg["cap"], cfg["frames"]
    thr = np.asarray(KEYPOINT_THRES_LIST, np.float32)
    hts = np.asarray(HEIGHTS, np.float64)

    # ---- parameters of the chained call (the same struct for the device-resident and the host form)
    def params(mem):
        prm = _capi.DetPoseParams()
        prm.n_frames, prm.h, prm.w = F, FRAME_H, FRAME_W
        prm.persons_cap, prm.swap_rb, prm.quarter_offset, prm.mem = cap, 0, 0, mem
        if cfg["detector"] == "yolo":
            prm.detector, prm.person_class = _capi.DET_YOLO, 0
            prm.N, prm.nc, prm.in_h, prm.in_w = dets[0].shape[1], dets[0].shape[2] - 5, 640, 640
            prm.letterbox_mode, prm.max_det, prm.cand_cap, prm.conf_thres, prm.iou_thres = 0, 300, 4096, 0.4, 0.5
        else:
            prm.detector, prm.person_class = _capi.DET_EDET, 1
            prm.K, prm.max_persons, prm.det_thres, prm.x_expand, prm.y_expand = 100, 16, 0.70, float(FRAME_H // 17), 0.0
        return prm

    # ---- device-resident inputs for `value`
    d_frames = eng.to_device(frames)
    d_dets = [eng.to_device(d) for d in dets] + [None, None]
    prm_dev = params(DEVICE)

    def step_device(time_hrnet=False):
        tk = C.c_int(-1)
        check(lib.hbp_det_pose_submit(ctx, C.byref(prm_dev), C.c_void_p(d_frames), ptr(d_dets[0]), ptr(d_dets[1]), ptr(d_dets[2]),
                                      ptr(hts), hts.size, ptr(thr), C.byref(tk)))
        return tk.value

    n_out, st_out = C.c_int(0), C.c_int(0)

    def collect(tk):
        check(lib.hbp_det_pose_collect(ctx, tk, C.byref(n_out), C.byref(st_out), None, None, None, None, None, None, None, None))
        return n_out.value

    for _ in range(max(args.warmup, 3)):          # >= 3 warm-up steps; the 2nd captures the HRNet CUDA graph
        n_live = collect(step_device())
    eng.sync()
    assert st_out.value == 0, "pipeline status %d" % st_out.value

    # ---- `value`: whole-job throughput, inputs resident in HBM.  The streaming form of the product: HBP_E2E_CONTEXTS
    # (default 2) engine contexts on this rank's GPU, steps alternating between them, two tickets in flight per context (the
    # forward of one step overlaps the tail of the previous one).  The steps cycle through input sets that together
    # exceed the 126 MB L2, so no step finds its frame or detector head cached.
    n_ctx = max(1, int(os.environ.get("HBP_E2E_CONTEXTS", "2")))
    pool_engines = [eng]
    for _ in range(n_ctx - 1):
        e2 = Engine(local_rank)
        e2.load_hrnet(hr_weights, cfg["width"], cfg["in_h"], cfg["in_w"], seed=0)
        pool_engines.append(e2)
    set_bytes = frames.nbytes + sum(d.nbytes for d in dets)
    n_sets = max(2, int(140e6 // set_bytes) + 2)
    d_sets = [(d_frames, d_dets)] + [(eng.to_device(frames), [eng.to_device(d) for d in dets] + [None, None]) for _ in range(n_sets - 1)]

    def dev_submit(i):
        e = pool_engines[i % n_ctx]
        fr, dd = d_sets[i % n_sets]
        tk = C.c_int(-1)
        check(lib.hbp_det_pose_submit(e._ctx, C.byref(prm_dev), C.c_void_p(fr), ptr(dd[0]), ptr(dd[1]), ptr(dd[2]),
                                      ptr(hts), hts.size, ptr(thr), C.byref(tk)))
        return e, tk.value

    def dev_collect(e, tk):
        check(lib.hbp_det_pose_collect(e._ctx, tk, C.byref(n_out), C.byref(st_out), None, None, None, None, None, None, None, None))
        assert st_out.value == 0, "pipeline status %d" % st_out.value
        return n_out.value

    def dev_stream(n_steps):
        inflight, crops = [], 0
        for i in range(n_steps):
            inflight.append(dev_submit(i))
            if len(inflight) >= 2 * n_ctx:
                crops += dev_collect(*inflight.pop(0))
        for it in inflight:
            crops += dev_collect(*it)
        return crops

    dev_stream(max(args.warmup, 3) * n_ctx)         # >= 3 warm-up steps per context (plans, graphs)
    for e in pool_engines:
        e.sync()
    sampler = ClockSampler(local_rank, period_ms=20) if rank == 0 else None
    barrier()
    if sampler:
        sampler.mark()
    launches0 = sum(e.kernel_launches() for e in pool_engines)
    t_wall0 = time.perf_counter()
    crops_done = dev_stream(args.steps)
    for e in pool_engines:
        e.sync()
    barrier()
    wall_ms = (time.perf_counter() - t_wall0) * 1e3
    launches = sum(e.kernel_launches() for e in pool_engines) - launches0
    clocks = sampler.stop() if sampler else None

    # ---- one context, one step at a time (CUDA events around each step, L2 flushed before it): the per-step device time
    # that the stage table and the roofline refer to
    step_ms = []
    single_crops = 0
    for _ in range(args.steps):
        eng.flush_l2()
        eng.timer_start(0)
        tk = step_device()
        eng.timer_stop(0)
        single_crops += collect(tk)
        step_ms.append(eng.timer_ms(0))
    eng.sync()
    dev_ms_total = sum(step_ms)

    # HRNet stage alone (CUDA events around hbp_hrnet_forward on the pipeline's own crop buffer shapes), for the roofline
    d_crops = eng.dev_alloc(cap * 3 * cfg["in_h"] * cfg["in_w"] * 2)
    d_hm = eng.dev_alloc(cap * 17 * (cfg["in_h"] // 4) * (cfg["in_w"] // 4) * 2)
    check(lib.hbp_memset_dev(ctx, C.c_void_p(d_crops), 0, cap * 3 * cfg["in_h"] * cfg["in_w"] * 2))
    hrnet_ms = []
    for i in range(3 + min(args.steps, 10)):
        eng.flush_l2()
        eng.timer_start(1)
        check(lib.hbp_hrnet_forward(ctx, C.c_void_p(d_crops), cap, C.c_void_p(d_hm), F16, DEVICE))
        eng.timer_stop(1)
        if os.environ.get("HBP_BENCH_DEBUG"):
            try:
                eng.sync()
            except Exception:
                sys.stderr.write("[bench debug] HRNet-only forward %d (0 = eager, 1 = graph capture, 2.. = replays) faulted\n" % i)
                raise
        if i >= 3:
            hrnet_ms.append(eng.timer_ms(1))
    eng.sync()                                     # (a device fault of this phase is reported here, not in the e2e phase below)
    # the same forward on every context of the pool at once (what the streaming form does to the conv stack)
    hr_pool_ms = None
    if n_ctx > 1:
        bufs = [(d_crops, d_hm)]
        for e2 in pool_engines[1:]:
            dc = e2.dev_alloc(cap * 3 * cfg["in_h"] * cfg["in_w"] * 2)
            check(lib.hbp_memset_dev(e2._ctx, C.c_void_p(dc), 0, cap * 3 * cfg["in_h"] * cfg["in_w"] * 2))
            bufs.append((dc, e2.dev_alloc(cap * 17 * (cfg["in_h"] // 4) * (cfg["in_w"] // 4) * 2)))
        reps = 3 + min(args.steps, 10)
        for i in range(3 + reps):
            if i == 3:
                for e in pool_engines:
                    e.sync()
                t_p = time.perf_counter()
            for e, (dc, dh) in zip(pool_engines, bufs):
                check(lib.hbp_hrnet_forward(e._ctx, C.c_void_p(dc), cap, C.c_void_p(dh), F16, DEVICE))
        for e in pool_engines:
            e.sync()
        hr_pool_ms = (time.perf_counter() - t_p) * 1e3 / (reps * n_ctx)

    # ---- e2e through the public API with host buffers (pinned), two batches in flight
    h_frames = [eng.pinned_empty(frames.shape, np.uint8) for _ in range(2)]
    h_dets = [[eng.pinned_empty(d.shape, d.dtype) for d in dets] for _ in range(2)]
    for k in range(2):
        h_frames[k][...] = frames
        for a, b in zip(h_dets[k], dets):
            a[...] = b

    def submit_host(k):
        if cfg["detector"] == "yolo":
            return eng.det_pose_submit_yolo(h_frames[k], h_dets[k][0], person_height=HEIGHTS, persons_cap=cap, resample="bilinear")
        return eng.det_pose_submit_edet(h_frames[k], h_dets[k][0], h_dets[k][1], h_dets[k][2], person_height=HEIGHTS,
                                        persons_cap=cap, max_persons=16)

    for _ in range(3):
        out = eng.det_pose_collect(submit_host(0))
    barrier()
    lat = []                                       # per-batch latency: one synchronous call at a time, nothing overlapped
    for _ in range(min(args.steps, 10)):
        t1 = time.perf_counter()
        out = eng.det_pose_collect(submit_host(0))
        lat.append((time.perf_counter() - t1) * 1e3)
    e2e_steps = args.steps if args.steps < 5 else max(args.steps, 50)
    # Throughput: the public streaming form -- MultiGpuEngine.det_pose_stream over HBP_E2E_CONTEXTS (default 2) engine
    # contexts on this rank's GPU, two tickets in flight each, steps alternating between them.  Every step still uploads its
    # frame + detector head from pinned memory and reads its results back.
    from human_body_proportion_estimation_b200.engine import MultiGpuEngine
    pool = MultiGpuEngine(engines=pool_engines)
    pool_bufs = [(h_frames, h_dets)]
    for e2 in pool_engines[1:]:
        hf = [e2.pinned_empty(frames.shape, np.uint8) for _ in range(2)]
        hd = [[e2.pinned_empty(d.shape, d.dtype) for d in dets] for _ in range(2)]
        for k in range(2):
            hf[k][...] = frames
            for a, b in zip(hd[k], dets):
                a[...] = b
        pool_bufs.append((hf, hd))

    def pool_submit(e, r, s):
        hf, hd = pool_bufs[r]
        k = s & 1
        if cfg["detector"] == "yolo":
            return e.det_pose_submit_yolo(hf[k], hd[k][0], person_height=HEIGHTS, persons_cap=cap, resample="bilinear")
        return e.det_pose_submit_edet(hf[k], hd[k][0], hd[k][1], hd[k][2], person_height=HEIGHTS, persons_cap=cap, max_persons=16)

    pool.det_pose_stream(pool_submit, 4 * n_ctx)              # warm-up of the additional contexts (plans, graphs)
    sampler2 = ClockSampler(local_rank, period_ms=250) if rank == 0 else None
    barrier()
    if sampler2:
        sampler2.mark()
    t0 = time.perf_counter()
    res = pool.det_pose_stream(pool_submit, e2e_steps)
    e2e_s = time.perf_counter() - t0
    e2e_crops = sum(r["n"] for r in res)
    out = eng.det_pose_collect(submit_host(0), return_heatmaps=True)      # (checked against the oracle below)
    barrier()
    h2d = frames.nbytes + sum(d.nbytes for d in dets) + hts.nbytes + 17 * 4 + 64
    d2h = 64 + cap * (8 + 4 + 16 + 17 * 8 + 17 * 4 + 4 + 44)
    if sampler2:
        c2 = sampler2.stop()
        if clocks and c2.get("samples"):
            clocks["e2e_region"] = c2
            clocks["reasons"] = sorted(set(clocks.get("reasons", [])) | set(c2.get("reasons", [])))

    # ---- max over ranks (device time, e2e wall time)
    dev_ms_total, e2e_s, wall_ms = grp.max_over_ranks([dev_ms_total, e2e_s, wall_ms])
    launches = int(grp.sum_over_ranks([launches])[0])
    crops_all = grp.sum_over_ranks([crops_done, e2e_crops, single_crops])
    if rank != 0:
        grp.close()
        return

    parity = check_against_oracle(out, HEIGHTS)

    # ---- per-stage timings on the detector-head shapes, device-resident (informational)
    stages = {}
    try:
        from human_body_proportion_estimation_b200 import synth
        pred = synth.yolo_decoded_head()[0]
        d_pred = eng.to_device(pred)
        d_det = eng.dev_alloc(300 * 6 * 4)
        d_cnt = eng.dev_alloc(4)
        d_cls = eng.to_device(np.zeros(1, np.int32))
        d_lb = eng.dev_alloc(3 * 640 * 640 * 2)
        d_frame1 = d_frames                                   # first frame

        def timed(fn, reps=20):
            fn(); eng.sync()
            eng.timer_start(2)
            for _ in range(reps):
                fn()
            eng.timer_stop(2)
            return eng.timer_ms(2) / reps

        stages["letterbox_1080p_to_640_f16"] = timed(lambda: check(lib.hbp_preprocess(
            ctx, C.c_void_p(d_frame1), 1, FRAME_H, FRAME_W, PRE_LETTERBOX, 640, 640, 1, 128, C.c_void_p(d_lb), F16, NCHW, DEVICE)))
        stages["letterbox_pil_bicubic_1080p_to_640_f16"] = timed(lambda: check(lib.hbp_preprocess(
            ctx, C.c_void_p(d_frame1), 1, FRAME_H, FRAME_W, PRE_LETTERBOX_PIL, 640, 640, 1, 128, C.c_void_p(d_lb), F16, NCHW, DEVICE)))
        stages["yolo_nms_25200x85_person"] = timed(lambda: check(lib.hbp_yolo_nms(
            ctx, C.c_void_p(d_pred), 1, 25200, 80, 0.4, 0.5, C.c_void_p(d_cls), 1, 300, C.c_void_p(d_det), C.c_void_p(d_cnt), DEVICE)))
        stages["hrnet_w%d_%dslots" % (cfg["width"], cap)] = statistics.mean(hrnet_ms)
        stages["chain_total"] = dev_ms_total / args.steps
        stages["chain_minus_hrnet (letterbox + NMS + person params + crop + decode + result copy)"] = dev_ms_total / args.steps - statistics.mean(hrnet_ms)
    except Exception as e:           # stage timings are informational
        stages["error"] = str(e)

    pk = peaks()
    flops_crop, _ = hrnet_arch.flops_per_crop(cfg["width"], cfg["in_h"], cfg["in_w"])
    hr_ms = statistics.mean(hrnet_ms)
    achieved = flops_crop * cap / (hr_ms * 1e-3) / 1e12
    n_launch_step = max(1, int(launches) // max(1, world) // args.steps)
    value = crops_all[0] / (wall_ms * 1e-3)
    single_value = crops_all[2] / (dev_ms_total * 1e-3)
    e2e_val = crops_all[1] / e2e_s

    stage_rf = None
    cpu = None
    if world == 1:
        try:
            stage_rf = stage_rooflines(eng, pk["hbm"])
        except Exception as e:
            stage_rf = {"error": str(e)}
        try:          # the CPU baseline is timed on rank 0 at N = 1 only, one or two full batches of the same workload
            cv, cms, cores, note = cpu_reference_run(cfg, 2 if cfg["width"] == 32 else 1, 1, frames, dets)
            cpu = {"value": cv, "unit": UNIT, "cores": cores, "kind": "port", "sample": note}
        except Exception as e:
            cpu = {"value": None, "unit": UNIT, "cores": 0, "kind": "port", "sample": "failed: %s" % e}

    stream = None
    if args.config == 1:
        try:          # the product's one-process multi-GPU API on BASELINE configs[4] (4K stream, 100 persons per frame)
            stream = stream_config4(world)
        except Exception as e:
            stream = {"error": str(e)}

    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
        "warmup": max(args.warmup, 3), "ms_per_step": wall_ms / args.steps, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f16", "data": "synthetic",
        "config": config_dict(cfg),
        "crops_per_step_per_gpu": n_live,
        "contexts_per_gpu": n_ctx,
        "timing": "barrier + synchronize on both sides of exactly K steps; K steps / that time, max over ranks",
        "single_context": {"value": single_value, "unit": UNIT, "ms_per_step": dev_ms_total / args.steps,
                           "note": "one engine context, one step at a time, CUDA events around each step, L2 flushed before it"},
        "e2e": {"value": e2e_val, "unit": UNIT, "h2d_bytes_per_step": int(h2d), "d2h_bytes_per_step": int(d2h),
                "p50_batch_latency_ms": statistics.median(lat), "steps": e2e_steps,
                "contexts_per_gpu": n_ctx,
                "api": "MultiGpuEngine.det_pose_stream over %d engine context(s) on the GPU: Engine.det_pose_submit_%s / det_pose_collect (hbp_det_pose_submit/_collect), two tickets in flight per context; latency from one synchronous submit+collect at a time" % (n_ctx, cfg["detector"])},
        "parity_check": parity,
        "gpu_launches": int(launches),
        "roofline": {"bound": "tensor", "kernel": "HRNet conv stack (conv_umma_halo_kernel, conv_umma_pgroup_kernel, conv_umma_kernel, upsample_add_group, stem, head; one CUDA graph), %d person slots; the whole chain is %d launches per step" % (cap, n_launch_step),
                     "achieved": achieved, "peak": pk["tflops"], "unit": "TFLOP/s", "frac": achieved / pk["tflops"],
                     "peak_source": pk["src"], "traffic": hrnet_traffic().get("dram_bytes_per_launch"),
                     "traffic_note": hrnet_traffic().get("note"),
                     "flop_per_step": flops_crop * cap, "hrnet_ms": hr_ms,
                     "streaming": None if not hr_pool_ms else {
                         "note": "the same forward on %d engine contexts at once (host clock over %d forwards, no L2 flush): time per forward and the fraction of the peak it corresponds to" % (n_ctx, (3 + min(args.steps, 10)) * n_ctx),
                         "hrnet_ms_per_forward": hr_pool_ms, "achieved": flops_crop * cap / (hr_pool_ms * 1e-3) / 1e12,
                         "frac": flops_crop * cap / (hr_pool_ms * 1e-3) / 1e12 / pk["tflops"]}},
        "cpu_baseline": cpu,
        "stages_ms": stages,
        "stage_rooflines": stage_rf,
        "stream_config4": stream,
        "clocks": clocks,
    }
    emit(json.dumps(line))
    grp.close()


def stream_config4(n_gpus, frames=160, warmup=16, persons=100):
    """BASELINE configs[4] through MultiGpuEngine.stream in ONE process over `n_gpus` GPUs (frame f -> GPU f mod G, two
    frames in flight per GPU, one host thread per GPU)."""
    from human_body_proportion_estimation_b200 import geometry, synth
    from human_body_proportion_estimation_b200.engine import MultiGpuEngine
    H, W = 2160, 3840
    n_ctx = max(1, int(os.environ.get("HBP_E2E_CONTEXTS", "2")))
    mg = MultiGpuEngine([d for d in range(n_gpus) for _ in range(n_ctx)], width=32, in_h=256, in_w=192, seed=0)   # n_ctx engine contexts per GPU
    base = [synth.frame_u8(H, W, seed=synth.SEED_BASE + 50 + i, smooth=False) for i in range(4)]
    pinned = []
    for e in mg.engines:
        bufs = []
        for b in base:
            p = e.pinned_empty(b.shape, np.uint8)
            p[...] = b
            bufs.append(p)
        pinned.append(bufs)
    sets = []
    for i in range(4):
        boxes = synth.person_boxes_yxyx_px(persons, H, W, seed=synth.SEED_BASE + 60 + i, hmin=150, hmax=600)
        mats = geometry.crop_and_resize_matrices(boxes / np.array([H, W, H, W], np.float32), H, W, 256, 192)
        sets.append((mats.reshape(-1, 6), boxes))
    G = n_gpus * n_ctx                      # engines: frame f -> engine f mod G (GPU (f mod G) // n_ctx)

    def source(f):
        mats, boxes = sets[f % 4]
        return pinned[f % G][(f // G) % 4], mats, boxes, 175.0

    n = frames * n_gpus
    mg.stream(source, warmup * G)
    t0 = time.perf_counter()
    res, lat = mg.stream(source, n)
    dt = time.perf_counter() - t0
    # the synchronous latency: one frame at a time on GPU 0 (submit -> collect, nothing else in flight)
    e0 = mg.engines[0]
    sync_lat = []
    for f in range(8):
        fr, mats, boxes, h = source(f * G)
        t1 = time.perf_counter()
        e0.pose_pipeline_collect(e0.pose_pipeline_submit(fr, mats, np.zeros(persons, np.int32), boxes, h))
        sync_lat.append((time.perf_counter() - t1) * 1e3)
    lat = np.sort(np.asarray(lat))
    for e in mg.engines:
        e.close()
    return {"workload": "configs[4]: 4K frames (2160x3840x3 u8, pinned), %d persons/frame, HRNet-W32 256x192, ONE process, frame f -> engine f mod (GPUs x contexts per GPU)" % persons,
            "n_gpus": n_gpus, "contexts_per_gpu": n_ctx, "frames": n, "crops_per_s": n * persons / dt, "frames_per_s": n / dt,
            "p50_frame_latency_ms_two_in_flight": float(lat[len(lat) // 2]), "p95_frame_latency_ms_two_in_flight": float(lat[int(len(lat) * 0.95)]),
            "p50_frame_latency_ms_synchronous": float(np.median(sync_lat)),
            "api": "MultiGpuEngine.stream (hbp_pose_pipeline_submit/_collect)"}


_REAL_STDOUT = None


def emit(line):
    """the ONE JSON line of the contract, on the process's original stdout"""
    data = (line + "\n").encode()
    if _REAL_STDOUT is None:
        sys.stdout.write(line + "\n"); sys.stdout.flush()
    else:
        os.write(_REAL_STDOUT, data)


def main():
    # Libraries print to file descriptor 1 on their own (NCCL: "NCCL version ..." at communicator creation, more with
    # NCCL_DEBUG): everything but the JSON line goes to stderr
    global _REAL_STDOUT
    sys.stdout.flush()
    _REAL_STDOUT = os.dup(1)
    os.dup2(2, 1)
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--config", type=int, default=1, choices=[1, 2, 3])
    ap.add_argument("--retried", action="store_true", help=argparse.SUPPRESS)
    args = ap.parse_args()
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if args.impl == "reference":
        run_reference(args, rank)
        return
    if world == 1 and args.gpus > 1:
        # not under torchrun: relaunch one rank per GPU
        cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", str(args.gpus),
               "--master-addr", "127.0.0.1", "--master-port", str(29500 + os.getpid() % 1000), os.path.abspath(__file__),
               "--gpus", str(args.gpus), "--steps", str(args.steps), "--warmup", str(args.warmup), "--config", str(args.config)]
        sys.exit(subprocess.call(cmd))
    try:
        run_ours(args, rank, world, local_rank)
    except Exception:
        # a CUDA error is sticky for the process: say what happened and (single GPU, once) measure again in a fresh
        # process instead of leaving the caller without a line
        import traceback
        traceback.print_exc()
        sys.stderr.flush()
        if world > 1 or args.retried:
            raise
        os.dup2(_REAL_STDOUT, 1)                   # the fresh process gets the caller's stdout back
        os.execv(sys.executable, [sys.executable, os.path.abspath(__file__)] + sys.argv[1:] + ["--retried"])


if __name__ == "__main__":
    main()
