// Chained det -> pose pipeline (BASELINE configs[2] and [3]): frames + detector-head tensors in, per-person
// keypoints / scores / lengths out, with NO host round trip between the stages.
//
//   YOLO  (configs[2]):  letterbox (obj_det_yolov5_onnx.py:27-36) -> [detector backbone: not part of the reference tree,
//                        its decoded head (B,N,5+nc) is an input] -> non_max_suppression (onnx_utils.py:125-222, class
//                        filter) -> scale_coords (onnx_utils.py:252-266) -> per-person crop parameters -> crop ->
//                        HRNet -> decode + proportions
//   EDET  (configs[3]):  EfficientDet outputs (boxes, scores, classes) -> person filter + box expansion
//                        (models/conv.py:22-57) -> crop_and_resize parameters (:59-70) -> crop -> HRNet -> decode +
//                        proportions (person_det_pose_edet4_trtserver.py:145-171)
//
// The number of persons is only known on the device.  The pipeline therefore always computes `persons_cap` person slots
// (HRNet batch = persons_cap: one CUDA graph whatever the frame holds); the crop and decode kernels skip the slots
// beyond the device-side count, and the count comes back with the results.  status bit 0: more NMS candidates than
// cand_cap, bit 1: more persons than persons_cap (the surplus is dropped, the caller can re-run with larger caps).
#include "hbp_internal.cuh"
#include <cstring>

static int reserve(uint8_t** p, size_t* cap, size_t bytes, bool pinned) {
    if (bytes <= *cap) return HBP_OK;
    if (*p) { if (pinned) cudaFreeHost(*p); else cudaFree(*p); *p = nullptr; *cap = 0; }
    const size_t n = bytes + bytes / 4 + 256;
    cudaError_t e = pinned ? cudaMallocHost((void**)p, n) : cudaMalloc((void**)p, n);
    if (e != cudaSuccess) return hbp_cuda_fail(e, "det-pose slot allocation", __FILE__, __LINE__);
    *cap = n;
    return HBP_OK;
}

static size_t al(size_t x) { return (x + 63) & ~size_t(63); }

struct ResultLayout {        // one block per slot, device and pinned mirror
    size_t o_n, o_status, o_fidx, o_boxes, o_kpts, o_scores, o_ign, o_len, o_torso, bytes;
    explicit ResultLayout(int cap) {
        size_t o = 0;
        o_n = o; o += 64; o_status = o_n + 4;
        o_torso = o; o += al((size_t)cap * 8);
        o_fidx = o; o += al((size_t)cap * 4);
        o_boxes = o; o += al((size_t)cap * 16);
        o_kpts = o; o += al((size_t)cap * 17 * 8);
        o_scores = o; o += al((size_t)cap * 17 * 4);
        o_ign = o; o += al((size_t)cap * 4);
        o_len = o; o += al((size_t)cap * 44);
        bytes = o;
    }
};

extern "C" {

int hbp_det_pose_submit(hbp_ctx* ctx, const hbp_det_pose_params* prm, const uint8_t* frames, const float* det0,
                        const float* det1, const float* det2, const double* heights, int n_heights,
                        const float* joint_thr, int* ticket) {
    if (!ctx) { hbp_set_error("null context"); return HBP_ERR_INVALID; }
    HBP_CUDA(cudaSetDevice(ctx->device));
    HBP_REQUIRE(prm && frames && det0 && heights && n_heights > 0 && joint_thr && ticket, "null argument");
    HBP_REQUIRE(prm->n_frames > 0 && prm->n_frames <= 1024 && prm->h > 0 && prm->w > 0, "bad frame shape");
    HBP_REQUIRE(prm->persons_cap > 0, "persons_cap must be positive");
    HBP_REQUIRE(prm->detector == HBP_DET_YOLO || prm->detector == HBP_DET_EDET, "detector must be HBP_DET_YOLO or HBP_DET_EDET");
    if (!ctx->hrnet) { hbp_set_error("hbp_det_pose_submit before hbp_hrnet_load"); return HBP_ERR_STATE; }
    const int F = prm->n_frames, cap = prm->persons_cap, J = 17;
    if (prm->detector == HBP_DET_YOLO) {
        HBP_REQUIRE(prm->N > 0 && prm->nc > 0 && prm->in_h > 0 && prm->in_w > 0 && prm->max_det > 0, "bad YOLO head shape");
        HBP_REQUIRE(prm->cand_cap >= 32 && prm->cand_cap % 32 == 0, "cand_cap must be a positive multiple of 32");
        HBP_REQUIRE(prm->conf_thres >= 0.f && prm->conf_thres <= 1.f && prm->iou_thres >= 0.0 && prm->iou_thres <= 1.0, "bad thresholds");
    } else {
        HBP_REQUIRE(det1 && det2 && prm->K > 0 && prm->max_persons > 0, "bad EfficientDet output shape");
    }
    const int k = (int)(ctx->pipe_seq % HBP_PIPE_SLOTS);
    hbp_pipe_slot& sl = ctx->pipe[k];
    if (sl.busy) { hbp_set_error("pipeline slot %d has not been collected (at most %d batches in flight)", k, HBP_PIPE_SLOTS); return HBP_ERR_STATE; }
    if (!ctx->copy_stream) HBP_CUDA(cudaStreamCreateWithFlags(&ctx->copy_stream, cudaStreamNonBlocking));
    if (!sl.ev_h2d) {
        HBP_CUDA(cudaEventCreateWithFlags(&sl.ev_h2d, cudaEventDisableTiming));
        HBP_CUDA(cudaEventCreateWithFlags(&sl.ev_crop, cudaEventDisableTiming));
        HBP_CUDA(cudaEventCreateWithFlags(&sl.ev_done, cudaEventDisableTiming));
    }
    int ih, iw, wd;
    hrnet_dims(ctx, &ih, &iw, &wd);
    const int Hh = ih / 4, Wh = iw / 4;
    const size_t frame_bytes = (size_t)F * prm->h * prm->w * 3;
    // input block: detector tensors | heights | thresholds | class filter
    size_t d0 = 0, d1 = 0, d2 = 0;
    if (prm->detector == HBP_DET_YOLO) d0 = (size_t)F * prm->N * (5 + prm->nc) * 4;
    else { d0 = (size_t)F * prm->K * 16; d1 = (size_t)F * prm->K * 4; d2 = d1; }
    const size_t i_d0 = 0, i_d1 = i_d0 + al(d0), i_d2 = i_d1 + al(d1), i_h = i_d2 + al(d2), i_t = i_h + al((size_t)n_heights * 8),
                 i_c = i_t + al(32 * 4), in_bytes = i_c + 64;
    const ResultLayout R(cap);
    // device-only block: NMS / filter outputs, person parameters
    const int per_frame = prm->detector == HBP_DET_YOLO ? prm->max_det : prm->max_persons;
    const size_t w_det = 0, w_cnt = w_det + al((size_t)F * per_frame * 6 * 4), w_M = w_cnt + al((size_t)F * 4),
                 w_hcm = w_M + al((size_t)cap * 48), work_bytes = w_hcm + al((size_t)cap * 8);
    const size_t misc_bytes = al(in_bytes) + R.bytes + work_bytes;
    const bool dev_in = prm->mem == HBP_DEVICE;
    if ((!dev_in && frame_bytes > sl.frames_cap) || misc_bytes > sl.misc_cap) HBP_CUDA(cudaDeviceSynchronize());
    int st = dev_in ? HBP_OK : reserve(&sl.d_frames, &sl.frames_cap, frame_bytes, false);
    if (!st) st = reserve(&sl.d_misc, &sl.misc_cap, misc_bytes, false);
    if (!st) st = reserve(&sl.h_pin, &sl.pin_cap, al(in_bytes) + R.bytes, true);
    if (st) return st;
    __half* d_crops = (__half*)hbp_scratch(ctx, SC_PIPE_CROPS, (size_t)cap * 3 * ih * iw * 2);
    void* d_hm = hbp_scratch(ctx, SC_PIPE_HM, (size_t)cap * J * Hh * Wh * 2);
    if (!d_crops || !d_hm) return HBP_ERR_NOMEM;
    uint8_t* d_in = sl.d_misc;
    uint8_t* d_res = sl.d_misc + al(in_bytes);
    uint8_t* d_work = d_res + R.bytes;
    uint8_t* h_in = sl.h_pin;
    // small inputs go through the slot's pinned mirror (the caller's arrays may be pageable); the detector tensors and the
    // frames are copied straight from the caller's buffers
    memcpy(h_in + i_h, heights, (size_t)n_heights * 8);
    memcpy(h_in + i_t, joint_thr, 17 * 4);
    const int person_class = prm->person_class;
    memcpy(h_in + i_c, &person_class, 4);
    cudaStream_t cs = ctx->copy_stream;
    if (sl.crop_recorded) HBP_CUDA(cudaStreamWaitEvent(cs, sl.ev_crop, 0));
    const uint8_t* d_frames = dev_in ? frames : sl.d_frames;
    const float* d_det0 = dev_in ? det0 : (const float*)(d_in + i_d0);
    const float* d_det1 = dev_in ? det1 : (const float*)(d_in + i_d1);
    const float* d_det2 = dev_in ? det2 : (const float*)(d_in + i_d2);
    if (!dev_in) {
        HBP_CUDA(cudaMemcpyAsync(sl.d_frames, frames, frame_bytes, cudaMemcpyHostToDevice, cs));
        HBP_CUDA(cudaMemcpyAsync(d_in + i_d0, det0, d0, cudaMemcpyHostToDevice, cs));
        if (d1) HBP_CUDA(cudaMemcpyAsync(d_in + i_d1, det1, d1, cudaMemcpyHostToDevice, cs));
        if (d2) HBP_CUDA(cudaMemcpyAsync(d_in + i_d2, det2, d2, cudaMemcpyHostToDevice, cs));
    }
    HBP_CUDA(cudaMemcpyAsync(d_in + i_h, h_in + i_h, in_bytes - i_h, cudaMemcpyHostToDevice, cs));
    HBP_CUDA(cudaEventRecord(sl.ev_h2d, cs));
    HBP_CUDA(cudaStreamWaitEvent(ctx->stream, sl.ev_h2d, 0));
    HBP_CUDA(cudaMemsetAsync(d_res, 0, 64, ctx->stream));                 // n_persons, status
    int* d_n = (int*)(d_res + R.o_n);
    int* d_status = (int*)(d_res + R.o_status);
    double* d_M = (double*)(d_work + w_M);
    double* d_hcm = (double*)(d_work + w_hcm);
    float* d_boxes = (float*)(d_res + R.o_boxes);
    int* d_fidx = (int*)(d_res + R.o_fidx);
    int s = HBP_OK;
    if (prm->detector == HBP_DET_YOLO) {
        // the tensor the detector backbone consumes (the reference builds it for every frame, obj_det_yolov5_onnx.py:107-113)
        void* d_lb = hbp_scratch(ctx, SC_PIPE_LB, (size_t)F * 3 * prm->in_h * prm->in_w * 2);
        if (!d_lb) return HBP_ERR_NOMEM;
        s = k_preprocess(ctx, d_frames, F, prm->h, prm->w, prm->letterbox_mode ? HBP_PRE_LETTERBOX_PIL : HBP_PRE_LETTERBOX,
                         prm->in_h, prm->in_w, prm->swap_rb, 128, d_lb, HBP_F16, HBP_NCHW);
        if (s) return s;
        float* d_det = (float*)(d_work + w_det);
        int* d_cnt = (int*)(d_work + w_cnt);
        s = k_yolo_nms_bounded(ctx, d_det0, F, prm->N, prm->nc, prm->conf_thres, prm->iou_thres,
                               person_class >= 0 ? (const int*)(d_in + i_c) : nullptr, person_class >= 0 ? 1 : 0, prm->max_det,
                               prm->cand_cap, d_det, d_cnt, d_status);
        if (s) return s;
        s = k_persons_from_yolo(ctx, d_det, d_cnt, F, prm->max_det, prm->in_h, prm->in_w, prm->h, prm->w, ih, iw,
                                (const double*)(d_in + i_h), n_heights, cap, d_M, d_boxes, d_fidx, d_hcm, d_n, d_status);
        if (s) return s;
    } else {
        float* d_fb = (float*)(d_work + w_det);
        int* d_cnt = (int*)(d_work + w_cnt);
        s = k_edet_filter(ctx, d_det0, d_det1, d_det2, F, prm->K,
                          (float)(person_class >= 0 ? person_class : 1), prm->det_thres, prm->x_expand, prm->y_expand, prm->h, prm->w,
                          prm->max_persons, d_fb, d_cnt);
        if (s) return s;
        s = k_persons_from_edet(ctx, d_fb, d_cnt, F, prm->max_persons, prm->h, prm->w, ih, iw, (const double*)(d_in + i_h),
                                n_heights, cap, d_M, d_boxes, d_fidx, d_hcm, d_n, d_status);
        if (s) return s;
    }
    s = k_crop_warp(ctx, d_frames, F, prm->h, prm->w, d_M, d_fidx, cap, ih, iw, prm->swap_rb, d_crops, HBP_F16, d_n);
    if (s) return s;
    HBP_CUDA(cudaEventRecord(sl.ev_crop, ctx->stream));
    sl.crop_recorded = true;
    s = hrnet_forward(ctx, d_crops, cap, d_hm, HBP_F16);
    if (s) return s;
    s = k_decode_proportions(ctx, d_hm, HBP_F16, cap, J, Hh, Wh, d_boxes, d_hcm, (const float*)(d_in + i_t), prm->quarter_offset,
                             nullptr, (float*)(d_res + R.o_kpts), (float*)(d_res + R.o_scores), nullptr,
                             (uint32_t*)(d_res + R.o_ign), (float*)(d_res + R.o_len), (double*)(d_res + R.o_torso), d_n);
    if (s) return s;
    HBP_CUDA(cudaMemcpyAsync(sl.h_pin + al(in_bytes), d_res, R.bytes, cudaMemcpyDeviceToHost, ctx->stream));
    HBP_CUDA(cudaEventRecord(sl.ev_done, ctx->stream));
    sl.busy = true; sl.P = cap; sl.par_bytes = al(in_bytes); sl.res_bytes = R.bytes; sl.det_pose = true;
    sl.hm_bytes = (size_t)cap * J * Hh * Wh * 2;
    *ticket = k;
    ctx->pipe_seq++;
    return HBP_OK;
}

int hbp_det_pose_collect(hbp_ctx* ctx, int ticket, int* n_persons, int* status, int* frame_idx, float* boxes_yxyx_px,
                         float* kpts_img, float* scores, uint32_t* ignored, float* lengths_cm, double* torso_cm,
                         void* heatmaps_f16) {
    if (!ctx) { hbp_set_error("null context"); return HBP_ERR_INVALID; }
    HBP_CUDA(cudaSetDevice(ctx->device));
    HBP_REQUIRE(ticket >= 0 && ticket < HBP_PIPE_SLOTS && n_persons, "bad ticket");
    hbp_pipe_slot& sl = ctx->pipe[ticket];
    if (!sl.busy || !sl.det_pose) { hbp_set_error("ticket %d has no det-pose batch in flight", ticket); return HBP_ERR_STATE; }
    HBP_CUDA(cudaEventSynchronize(sl.ev_done));
    const int cap = sl.P, J = 17;
    const ResultLayout R(cap);
    const uint8_t* h = sl.h_pin + sl.par_bytes;
    int n = 0;
    memcpy(&n, h + R.o_n, 4);
    if (n < 0) n = 0;
    if (n > cap) n = cap;
    *n_persons = n;
    if (status) memcpy(status, h + R.o_status, 4);
    if (frame_idx) memcpy(frame_idx, h + R.o_fidx, (size_t)n * 4);
    if (boxes_yxyx_px) memcpy(boxes_yxyx_px, h + R.o_boxes, (size_t)n * 16);
    if (kpts_img) memcpy(kpts_img, h + R.o_kpts, (size_t)n * J * 8);
    if (scores) memcpy(scores, h + R.o_scores, (size_t)n * J * 4);
    if (ignored) memcpy(ignored, h + R.o_ign, (size_t)n * 4);
    if (lengths_cm) memcpy(lengths_cm, h + R.o_len, (size_t)n * 44);
    if (torso_cm) memcpy(torso_cm, h + R.o_torso, (size_t)n * 8);
    if (heatmaps_f16 && n > 0) {
        // optional: the heatmaps of the live persons (they stay in the shared scratch until the next submit's HRNet runs)
        void* d_hm = ctx->scratch[SC_PIPE_HM];
        HBP_CUDA(cudaMemcpyAsync(heatmaps_f16, d_hm, sl.hm_bytes / cap * n, cudaMemcpyDeviceToHost, ctx->stream));
        HBP_CUDA(cudaStreamSynchronize(ctx->stream));
    }
    sl.busy = false; sl.det_pose = false;
    return HBP_OK;
}

}  // extern "C"
