/*
 * hbp.h -- C ABI of the B200-native top-down pose hot path.
 *
 * The reference (SamSamhuns/human_body_proportion_estimation) is pure Python
 * and has no FFI of its own: the seams this library drops in behind are the
 * Python call sites listed per function below (paths relative to the reference
 * root, human_body_length_est/ abbreviated hble/).  The Python package
 * human_body_proportion_estimation_b200/ binds these with ctypes and re-exposes
 * the reference's function names; INTEGRATION.md shows the binding a reference
 * maintainer would add.
 *
 * Conventions
 *   - every function returns 0 (HBP_OK) or a negative hbp_status; the message
 *     of the last failure on the calling thread is hbp_last_error().
 *   - one hbp_ctx per GPU, used by one host thread at a time.
 *   - `mem` selects where ALL array arguments of a call live:
 *       HBP_HOST   : host pointers (pageable or pinned).  The call copies
 *                    inputs to the device, runs, copies results back and
 *                    returns after they have landed (synchronous).
 *       HBP_DEVICE : device pointers.  The call only enqueues work on the
 *                    context's stream (asynchronous); use hbp_sync().
 *   - the caller owns every buffer it passes; the library owns its scratch,
 *     weights and CUDA graphs.
 *   - there is no CPU fallback anywhere: without a CUDA device
 *     hbp_ctx_create fails and nothing else can be called.
 */
#ifndef HBP_H_
#define HBP_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define HBP_VERSION 100

#if defined(HBP_BUILD) && defined(__GNUC__)
#define HBP_API __attribute__((visibility("default")))
#else
#define HBP_API
#endif

typedef struct hbp_ctx hbp_ctx;

typedef enum {
    HBP_OK = 0,
    HBP_ERR_INVALID = -1,   /* bad argument */
    HBP_ERR_CUDA = -2,      /* CUDA runtime/driver error (text in hbp_last_error) */
    HBP_ERR_NOMEM = -3,
    HBP_ERR_STATE = -4,     /* e.g. forward before load */
    HBP_ERR_OVERFLOW = -5   /* result does not fit the caller's buffer */
} hbp_status;

typedef enum { HBP_HOST = 0, HBP_DEVICE = 1 } hbp_mem;
typedef enum { HBP_U8 = 0, HBP_F16 = 1, HBP_F32 = 2 } hbp_dtype;
typedef enum { HBP_NCHW = 0, HBP_NHWC = 1 } hbp_layout;

/* ---- context, memory, timing ------------------------------------------- */
HBP_API int hbp_version(void);
HBP_API const char* hbp_last_error(void);                 /* thread-local */
HBP_API int hbp_device_count(int* n);
HBP_API int hbp_ctx_create(int device, hbp_ctx** out);
HBP_API int hbp_ctx_destroy(hbp_ctx* ctx);
HBP_API int hbp_sync(hbp_ctx* ctx);
HBP_API int hbp_dev_alloc(hbp_ctx* ctx, size_t nbytes, void** out);
HBP_API int hbp_dev_free(hbp_ctx* ctx, void* p);
HBP_API int hbp_host_alloc(hbp_ctx* ctx, size_t nbytes, void** out);   /* pinned */
HBP_API int hbp_host_free(hbp_ctx* ctx, void* p);
HBP_API int hbp_copy_h2d(hbp_ctx* ctx, void* dst_dev, const void* src_host, size_t nbytes);  /* async on ctx stream */
HBP_API int hbp_copy_d2h(hbp_ctx* ctx, void* dst_host, const void* src_dev, size_t nbytes);  /* async on ctx stream */
HBP_API int hbp_memset_dev(hbp_ctx* ctx, void* dst_dev, int byte, size_t nbytes);
/* CUDA-event stopwatch on the context's stream (the stream every kernel of
 * this library is launched on).  Up to 8 independent slots. */
HBP_API int hbp_timer_start(hbp_ctx* ctx, int slot);
HBP_API int hbp_timer_stop(hbp_ctx* ctx, int slot);          /* records the stop event */
HBP_API int hbp_timer_elapsed_ms(hbp_ctx* ctx, int slot, float* ms); /* syncs on the stop event */
HBP_API int hbp_flush_l2(hbp_ctx* ctx);                       /* writes a 256 MiB scratch buffer */
HBP_API int hbp_kernel_launches(hbp_ctx* ctx, uint64_t* n);   /* kernels of this library launched so far
                                                         (graph replays count their nodes) */

/* ---- K1: frame preprocessing ------------------------------------------- *
 * Replaces hble/person_det_pose_edet4_trtserver.py:15-18 (mode COPY),
 * hble/modules/pose_estimator.py:29-45 and hble/pose_est_hrnet_trtserver.py:15-19
 * (mode STRETCH: cv2.resize-exact 11-bit bilinear to out_w x out_h),
 * hble/obj_det_yolov5_onnx.py:27-36 + hble/modules/onnx_utils.py:225-235
 * (mode LETTERBOX_PIL: the reference's scale/int()/centred-paste geometry on grey
 * `pad_value` with the reference's own resampler, PIL's antialiased BICUBIC
 * (hble/modules/onnx_utils.py:232), reproduced bit for bit; mode LETTERBOX: same
 * geometry with the cv2.resize bilinear sampler).
 * frames: (n,h,w,3) u8.  swap_rb!=0 reverses the channel order (BGR<->RGB).
 * out: (n,3,out_h,out_w) or (n,out_h,out_w,3); HBP_U8 keeps 0..255, HBP_F16 /
 * HBP_F32 store value/255 (correctly rounded).  In COPY mode out_h/out_w must
 * equal h/w. */
typedef enum { HBP_PRE_COPY = 0, HBP_PRE_STRETCH = 1, HBP_PRE_LETTERBOX = 2, HBP_PRE_LETTERBOX_PIL = 3 } hbp_pre_mode;
HBP_API int hbp_preprocess(hbp_ctx* ctx, const uint8_t* frames, int n, int h, int w,
                   int mode, int out_h, int out_w, int swap_rb, int pad_value,
                   void* out, int out_dtype, int out_layout, int mem);

/* ---- K2: YOLOv5 raw-head decode ---------------------------------------- *
 * Replaces hble/obj_det_yolov5_onnx.py:123-169.  heads[l]: (B,3,S_l,S_l,5+nc)
 * f32 in output order (anchors indexed by that order, :130-131).
 * out: (B, sum 3*S_l^2, 5+nc) f32 rows [cx,cy,w,h,obj,cls...]. */
HBP_API int hbp_yolo_decode_raw(hbp_ctx* ctx, const float* head0, const float* head1, const float* head2,
                        int B, int s0, int s1, int s2, int nc, int in_w, int in_h,
                        float* out, int mem);

/* ---- K2+K3: candidate filter + class-offset bitmask NMS ------------------ *
 * Replaces hble/modules/onnx_utils.py:125-222 (non_max_suppression; best-class,
 * non-agnostic, no merge -- the configuration every reference caller uses) and
 * the torchvision.ops.nms call at :205.
 * pred: (B,N,5+nc) f32.  classes: optional class-id filter (n_classes may be 0).
 * out_det: (B,max_det,6) f32 rows [x1,y1,x2,y2,conf,cls] score-descending,
 * out_count: (B) int32.  iou_thres is a double because torchvision compares the
 * float32 ratio against a double. */
HBP_API int hbp_yolo_nms(hbp_ctx* ctx, const float* pred, int B, int N, int nc,
                 float conf_thres, double iou_thres, const int* classes, int n_classes,
                 int max_det, float* out_det, int* out_count, int mem);

/* K2 alone: the candidate filter of non_max_suppression (hble/modules/onnx_utils.py:133,168-187: obj > conf, best
 * class of cls*obj > conf, optional class filter) on pred (B,N,5+nc) f32.  out_count (B) int32 = candidates per
 * image (capped at cand_cap); the candidates themselves stay in the library's scratch for hbp_yolo_nms-style
 * post-processing.  The stage's roofline probe (bench.py). */
HBP_API int hbp_yolo_filter(hbp_ctx* ctx, const float* pred, int B, int N, int nc, float conf_thres,
                    const int* classes, int n_classes, int cand_cap, int* out_count, int mem);

/* Legacy per-class greedy NMS with the +1 pixel IoU: replaces
 * hble/modules/onnx_utils.py:39-95 (w_non_max_suppression + w_bbox_iou).
 * out_det: (B,max_out,7) rows [x1,y1,x2,y2,obj,cls_conf,cls] grouped by class
 * ascending, obj-descending inside a class; out_count[b] = -1 when image b had
 * no candidate (the reference leaves None).  The reference's in-place rewrite
 * of pred[..., :4] to corners (:47) is done by the Python wrapper.
 * Rows are EXACTLY 5 + nc floats wide (row stride = 5 + nc): a head with more columns than the classes the caller wants
 * (the reference slices [:, 5:5+num_classes], :59) must be sliced to that width first -- Engine.yolo_nms_legacy does. */
HBP_API int hbp_yolo_nms_legacy(hbp_ctx* ctx, const float* pred, int B, int N, int nc,
                        float conf_thres, float nms_thres, int max_out,
                        float* out_det, int* out_count, int mem);

/* hble/modules/onnx_utils.py:238-266 (scale_coords + clip_coords), in place on
 * boxes (n,4) xyxy float32. */
HBP_API int hbp_scale_coords(hbp_ctx* ctx, float* boxes, int n, int img1_h, int img1_w,
                     int img0_h, int img0_w, int mem);

/* ---- K3b: EfficientDet person filter ------------------------------------ *
 * Replaces models/conv.py:22-57.  boxes (F,K,4) yxyx px, scores (F,K),
 * classes (F,K) (K=100 in the reference).  Keeps class==person_class and
 * score>=det_thres in detector order, at most max_persons (reference: 3),
 * expands by (x_expand,y_expand) px, clips to the frame, divides by [h,w,h,w].
 * out_boxes: (F,max_persons,4) yxyx normalised, out_count: (F). */
HBP_API int hbp_edet_person_filter(hbp_ctx* ctx, const float* boxes, const float* scores,
                           const float* classes, int F, int K, float person_class,
                           float det_thres, float x_expand, float y_expand,
                           int img_h, int img_w, int max_persons,
                           float* out_boxes, int* out_count, int mem);

/* ---- K4: per-person crop (batched affine warp) --------------------------- *
 * Replaces the crop of models/conv.py:59-80 and hble/modules/pose_estimator.py:29-45.
 * frames: (n_frames,h,w,3) u8.  M: (P,6) DOUBLE row-major 2x3 dst->src matrices
 * (what cv2.warpAffine uses with WARP_INVERSE_MAP); frame_idx: (P) int32.
 * Sampling is cv2.warpAffine(INTER_LINEAR, BORDER_CONSTANT 0)'s fixed point
 * (1/1024 px coordinates, 1/32 px weights) bit for bit; the result is divided
 * by 255 and stored as (P,3,out_h,out_w) fp16 (HBP_F16) or fp32 (HBP_F32). */
HBP_API int hbp_crop_warp(hbp_ctx* ctx, const uint8_t* frames, int n_frames, int h, int w,
                  const double* M, const int* frame_idx, int P, int out_h, int out_w,
                  int swap_rb, void* out, int out_dtype, int mem);

/* ---- K5: HRNet ----------------------------------------------------------- *
 * Replaces the opaque network behind hble/modules/pose_estimator.py:47-59
 * (onnxruntime) and the Triton `hrnet` model of the ensemble
 * (hble/person_det_pose_edet4_trtserver.py:22-23).  Weights: BN-folded fp16
 * blob in the layer order of hbp_hrnet_describe (below; the Python package's
 * hrnet_arch.py fills it).  width 32|48; in_h,in_w multiples of 32 (256x192, 384x288). */
HBP_API int hbp_hrnet_load(hbp_ctx* ctx, int width, int in_h, int in_w,
                   const void* weights_f16, size_t n_weight_halfs,
                   const float* biases_f32, size_t n_biases);
/* Host-only (no context, no GPU): the conv program of an architecture as text,
 * one line per convolution in weight-blob order:
 *   "<public HRNet state_dict prefix> <cin> <cout> <k> <stride> <w_off> <b_off> <out_h> <out_w> <up>\n"
 * (out_h,out_w = conv output size before the fused nearest upsample x<up>)
 * weights of a conv are [tap][cout][cin] halfs at w_off, biases floats at b_off.
 * *needed = bytes required for buf (incl. NUL). */
HBP_API int hbp_hrnet_describe(int width, int in_h, int in_w, char* buf, size_t buf_bytes,
                               size_t* n_weights, size_t* n_biases, size_t* needed);
/* crops: (P,3,in_h,in_w) fp16 NCHW in [0,1].  heatmaps: (P,17,in_h/4,in_w/4),
 * HBP_F16 or HBP_F32. */
HBP_API int hbp_hrnet_forward(hbp_ctx* ctx, const void* crops_f16, int P, void* heatmaps,
                      int out_dtype, int mem);
/* One fused convolution of the HRNet program on caller buffers (the operator the
 * network is made of):  out = act(conv_{k,stride}(in) + bias [+ residual]), NHWC
 * fp16, weights [tap][cout][cin] fp16, zero padding k/2, optional nearest
 * upsample x`up` folded into the store (residual read at the upsampled position).
 * engine 0 = SIMT tiles, 1 = tcgen05/TMA (falls back to 0 for shapes it does not
 * cover; *used_engine reports which one ran). */
HBP_API int hbp_conv2d_nhwc(hbp_ctx* ctx, int engine, const void* in_f16, int P, int H, int W, int Cin,
                            const void* weights_f16, const float* bias, const void* residual_f16,
                            int Cout, int k, int stride, int up, int relu, void* out_f16,
                            int* used_engine, int mem);
/* Same operator, launched `iters` more times back to back between two CUDA events on the
 * context's stream: *avg_ms = device time per launch (the kernel-level roofline probe). */
HBP_API int hbp_conv2d_nhwc_timed(hbp_ctx* ctx, int engine, const void* in_f16, int P, int H, int W, int Cin,
                                  const void* weights_f16, const float* bias, const void* residual_f16,
                                  int Cout, int k, int stride, int up, int relu, void* out_f16,
                                  int* used_engine, int mem, int iters, float* avg_ms);
/* which conv engine the loaded model runs: 0 = SIMT direct conv,
 * 1 = tcgen05/TMEM implicit GEMM fed by TMA. */
HBP_API int hbp_hrnet_set_engine(hbp_ctx* ctx, int engine);
/* debug/parity hooks (tests/test_gpu_hrnet_parity.py: per-stage error attribution against the fp32 oracle).
 * hbp_hrnet_debug_tensor copies the output tensor of program op `op_index` (NHWC fp16, channels as stored:
 * W48's 48/96-channel tensors are padded to 64/128) to the host; buffers are reused along the program, so a
 * tensor is only intact while it is live -- hbp_hrnet_forward_until runs the program eagerly (same plans and
 * kernels as the CUDA-graph path) up to and including op `op_index` and stops there.  hbp_hrnet_op_name
 * enumerates the program: *n_ops = number of ops, buf = name of op `op_index` (public HRNet state_dict prefixes
 * for convolutions; "<module>.fuse_levelK" / "<module>.fuse_upaddK" / "<module>.fuse_layers.I.upadd" for the
 * grouped launches and the upsample-adds). */
HBP_API int hbp_hrnet_debug_tensor(hbp_ctx* ctx, int op_index, void* out_host, size_t max_bytes,
                           int* n, int* h, int* w, int* c);
HBP_API int hbp_hrnet_forward_until(hbp_ctx* ctx, const void* crops_f16, int P, int op_index, int mem);
HBP_API int hbp_hrnet_op_name(hbp_ctx* ctx, int op_index, char* buf, size_t buf_bytes, int* n_ops);

/* ---- K6: heatmap decode fused with body-proportion geometry -------------- *
 * Replaces hble/modules/pose_estimator.py:74-99 (argmax decode),
 * hble/person_det_pose_edet4_trtserver.py:145-168 (remap, per-joint gate,
 * pixel->cm) and hble/modules/pose_estimator.py:130-200 (11 segment lengths).
 * heatmaps: (P,J,Hh,Wh) HBP_F32 or HBP_F16.  boxes_yxyx_px (P,4) float32 (the
 * already de-normalised boxes), height_cm (P) double, joint_thr (J) float32:
 * pass boxes == NULL to decode only.  quarter_offset != 0 adds the public
 * HRNet +-0.25 px step (not in the reference).
 * Outputs (any may be NULL): kpts_hm (P,J,2) heatmap-space (x,y); kpts_img
 * (P,J,2) image px; scores (P,J); argmax_idx (P,J) int32; ignored (P) bit j =
 * joint j below its threshold; lengths_cm (P,11) float32 in the order
 * shoulder,torso,lshoulder_lelbow,rshoulder_relbow,lwrist_lelbow,rwrist_relbow,
 * rhip_lhip,rhip_rknee,lhip_lknee,rankle_rknee,lankle_lknee with 0 = "Part not
 * visible"; torso_cm (P) double = the torso entry before rounding to float32
 * (the reference's torso is float64).  Lengths need J == 17. */
HBP_API int hbp_decode_proportions(hbp_ctx* ctx, const void* heatmaps, int dtype, int P, int J,
                           int Hh, int Wh, const float* boxes_yxyx_px,
                           const double* height_cm, const float* joint_thr,
                           int quarter_offset, float* kpts_hm, float* kpts_img,
                           float* scores, int32_t* argmax_idx, uint32_t* ignored,
                           float* lengths_cm, double* torso_cm, int mem);

/* ---- fused pipeline: frames + boxes -> crops -> HRNet -> decode ----------- *
 * The in-process replacement of the Triton ensemble call
 * (hble/modules/triton_utils.py:163-171) plus the per-person loop
 * (hble/person_det_pose_edet4_trtserver.py:148-171) for a batch of frames whose
 * person boxes are known (from hbp_edet_person_filter / hbp_yolo_nms).
 * Host buffers in, host buffers out (HBP_HOST only); the heatmaps stay on the
 * device unless heatmaps_out != NULL. */
typedef struct {
    int n_frames, h, w;            /* frames (n_frames,h,w,3) u8 */
    int P;                         /* total persons over all frames */
    int swap_rb;
    int quarter_offset;
    int heatmap_dtype;             /* HBP_F16 | HBP_F32 for heatmaps_out */
} hbp_pipeline_params;
HBP_API int hbp_pose_pipeline(hbp_ctx* ctx, const hbp_pipeline_params* prm, const uint8_t* frames,
                      const double* M, const int* frame_idx,
                      const float* boxes_yxyx_px, const double* height_cm,
                      const float* joint_thr,
                      float* kpts_img, float* scores, uint32_t* ignored,
                      float* lengths_cm, double* torso_cm, void* heatmaps_out);

/* Asynchronous form of hbp_pose_pipeline: up to two batches in flight.  submit() enqueues the
 * host->device copies on a copy stream and the kernels on the context's stream and returns a ticket;
 * collect() waits for that batch and copies the results out.  Calling submit(n+1) before collect(n)
 * overlaps the frame upload of n+1 with the network of n.  `frames` must stay valid (ideally
 * pinned, hbp_host_alloc) until the batch has been collected.  Same arrays as hbp_pose_pipeline;
 * heatmaps stay on the device. */
HBP_API int hbp_pose_pipeline_submit(hbp_ctx* ctx, const hbp_pipeline_params* prm, const uint8_t* frames,
                             const double* M, const int* frame_idx, const float* boxes,
                             const double* height_cm, const float* joint_thr, int* ticket);
HBP_API int hbp_pose_pipeline_collect(hbp_ctx* ctx, int ticket, float* kpts_img, float* scores,
                              uint32_t* ignored, float* lengths_cm, double* torso_cm);

/* ---- chained det -> pose pipeline (BASELINE configs[2], [3]) ---------------- *
 * Frames plus the detector's output tensors in, per-person results out, no host round trip between the stages:
 *   HBP_DET_YOLO: letterbox (hble/obj_det_yolov5_onnx.py:27-36) -> non_max_suppression on the decoded head
 *     (hble/modules/onnx_utils.py:125-222; det0 = (F,N,5+nc) f32, det1 = det2 = NULL; person_class >= 0 filters that
 *     class, < 0 keeps all) -> scale_coords (:252-266) -> int() box, cv2.resize-style crop
 *     (hble/modules/pose_estimator.py:29-45) -> HRNet -> decode + proportions.  Persons in score order per frame.
 *   HBP_DET_EDET: EfficientDet outputs det0 = boxes (F,K,4) yxyx px, det1 = scores (F,K), det2 = classes (F,K) ->
 *     person filter / expansion (models/conv.py:22-57) -> tf.image.crop_and_resize crop (:59-70) -> HRNet -> decode
 *     + proportions (hble/person_det_pose_edet4_trtserver.py:145-171).  Persons in detector order per frame.
 * The detector backbones are not part of the reference tree (README.md:13-26): their outputs are inputs here.
 * heights: person i of a frame gets heights[min(i, n_heights-1)] (hble/person_det_pose_edet4_trtserver.py:166-168).
 * Every call computes persons_cap person slots (one CUDA graph whatever a frame holds); *n_persons <= persons_cap
 * come back, in frame order.  status bit 0: an image had more NMS candidates than cand_cap; bit 1: more persons than
 * persons_cap (the surplus was dropped).  Same two-slot submit / collect protocol as hbp_pose_pipeline_submit. */
typedef enum { HBP_DET_YOLO = 0, HBP_DET_EDET = 1 } hbp_detector;
typedef struct {
    int n_frames, h, w;            /* frames (n_frames,h,w,3) u8 */
    int detector;                  /* hbp_detector */
    int persons_cap;
    int swap_rb, quarter_offset;
    int person_class;              /* YOLO: class id kept (reference callers: 0), EDET: class value kept (1) */
    /* YOLO */
    int N, nc, in_h, in_w;         /* head rows, classes, detector input size (640x640) */
    int letterbox_mode;            /* 0 = cv2-bilinear letterbox, 1 = the reference's PIL bicubic */
    int max_det, cand_cap;         /* kept boxes per frame (reference 300), NMS candidate capacity (multiple of 32) */
    float conf_thres;
    double iou_thres;
    /* EDET */
    int K, max_persons;            /* detector rows per frame (100), persons kept per frame (reference 3) */
    float det_thres, x_expand, y_expand;
    int mem;                       /* where frames / det0 / det1 / det2 live: HBP_HOST (copied in, the default) or HBP_DEVICE
                                      (used in place: the device-resident form bench.py times as `value`) */
} hbp_det_pose_params;
HBP_API int hbp_det_pose_submit(hbp_ctx* ctx, const hbp_det_pose_params* prm, const uint8_t* frames,
                        const float* det0, const float* det1, const float* det2,
                        const double* heights, int n_heights, const float* joint_thr, int* ticket);
/* any output may be NULL; arrays are sized for persons_cap rows, the first *n_persons are written.
 * heatmaps_f16: optional (persons_cap,17,Hh,Wh) fp16, valid until the next submit on this context. */
HBP_API int hbp_det_pose_collect(hbp_ctx* ctx, int ticket, int* n_persons, int* status, int* frame_idx,
                         float* boxes_yxyx_px, float* kpts_img, float* scores, uint32_t* ignored,
                         float* lengths_cm, double* torso_cm, void* heatmaps_f16);

/* ---- K6 with a general inverse affine (north_star item 5) -------------------- *
 * hbp_decode_proportions with one more input: M (P,6) double, the crop's dst->src matrices (what hbp_crop_warp
 * sampled with).  Keypoints are mapped back through them -- (u,v) = (x*crop_w/Wh, y*crop_h/Hh), image = M*(u,v,1) in
 * double, rounded once to float32 -- instead of the reference's box formula, so rotated / aspect-padded crops map
 * back exactly; boxes still provide pixel_to_cm = height_cm / (y2 - y1). */
HBP_API int hbp_decode_proportions_affine(hbp_ctx* ctx, const void* heatmaps, int dtype, int P, int J,
                           int Hh, int Wh, const float* boxes_yxyx_px, const double* M, int crop_h, int crop_w,
                           const double* height_cm, const float* joint_thr,
                           int quarter_offset, float* kpts_hm, float* kpts_img,
                           float* scores, int32_t* argmax_idx, uint32_t* ignored,
                           float* lengths_cm, double* torso_cm, int mem);

/* ---- segment lengths from keypoints the caller already holds ------------------- *
 * modules/pose_estimator.py:130-200 (get_keypoint_dist_dict) without the decode: kpts_img (P,17,2) float32 image px,
 * ignored (P) bit j = joint j ignored (NULL = none), pixel_to_cm (P) double -> lengths_cm (P,11) float32 in
 * HBP segment order (0 = not visible), torso_cm (P) double.  Same arithmetic as hbp_decode_proportions. */
HBP_API int hbp_keypoint_lengths(hbp_ctx* ctx, const float* kpts_img, const uint32_t* ignored,
                         const double* pixel_to_cm, int P, float* lengths_cm, double* torso_cm, int mem);

#ifdef __cplusplus
}
#endif
#endif /* HBP_H_ */
