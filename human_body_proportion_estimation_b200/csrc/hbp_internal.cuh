// Internal declarations shared by the translation units of libhbp_b200.so.
#pragma once
#include <cuda_runtime.h>
#include <cuda_fp16.h>
#include <stdint.h>
#include <stddef.h>
#include <string>
#include <vector>
#include "../../include/hbp.h"

struct HrnetModel;   // hrnet.cu

enum { HBP_SCRATCH_SLOTS = 24, HBP_TIMER_SLOTS = 8, HBP_PIPE_SLOTS = 2 };

// PIL-bicubic coefficient tables currently in SC_PRE_COEF (csrc/preprocess.cu)
struct hbp_pil_cache {
    bool valid = false;
    int W = 0, nw = 0, H = 0, nh = 0, ks_h = 0, ks_v = 0, row0 = 0, rows = 0;
    size_t o_kh = 0, o_bv = 0, o_kv = 0;
};

// one in-flight frame batch of the asynchronous pipeline (hbp_pose_pipeline_submit / _collect)
struct hbp_pipe_slot {
    uint8_t* d_frames = nullptr; size_t frames_cap = 0;
    uint8_t* d_misc = nullptr;   size_t misc_cap = 0;     // parameter block | result block
    uint8_t* h_pin = nullptr;    size_t pin_cap = 0;      // pinned mirror of d_misc
    cudaEvent_t ev_h2d = nullptr, ev_crop = nullptr, ev_done = nullptr;
    bool crop_recorded = false, busy = false;
    bool det_pose = false;                                // the batch in flight came from hbp_det_pose_submit
    int P = 0;
    size_t par_bytes = 0, res_bytes = 0, hm_bytes = 0;
};

struct hbp_ctx {
    int device = 0;
    int sm_count = 148;
    cudaStream_t stream = nullptr;
    cudaEvent_t ev_start[HBP_TIMER_SLOTS] = {};
    cudaEvent_t ev_stop[HBP_TIMER_SLOTS] = {};
    uint64_t launches = 0;
    uint32_t attr_flags = 0;     // per-device cudaFuncSetAttribute done (ATTR_*)
    // grow-only device scratch, one buffer per purpose so stages never alias
    void* scratch[HBP_SCRATCH_SLOTS] = {};
    size_t scratch_bytes[HBP_SCRATCH_SLOTS] = {};
    // pinned host staging for small result read-backs
    void* pinned = nullptr;
    size_t pinned_bytes = 0;
    void* l2_flush = nullptr;
    HrnetModel* hrnet = nullptr;
    hbp_pil_cache pil;
    cudaStream_t copy_stream = nullptr;      // host->device copies of the asynchronous pipeline
    hbp_pipe_slot pipe[HBP_PIPE_SLOTS];
    uint64_t pipe_seq = 0;
};

enum { ATTR_CROP = 1, ATTR_CONV = 2, ATTR_UMMA = 4, ATTR_NMS = 8, ATTR_PIL = 16, ATTR_RESIZE_U8 = 32, ATTR_RESIZE_F16 = 64, ATTR_RESIZE_F32 = 128, ATTR_FILTER = 256 };

// scratch slot ids
enum {
    SC_IN0 = 0, SC_IN1, SC_IN2, SC_IN3, SC_IN4, SC_IN5,   // host-mode input staging
    SC_OUT0, SC_OUT1, SC_OUT2, SC_OUT3, SC_OUT4, SC_OUT5, SC_OUT6,  // host-mode output staging
    SC_NMS_CAND, SC_NMS_SORTED, SC_NMS_MASK, SC_NMS_MISC,
    SC_PIPE_CROPS, SC_PIPE_HM, SC_PIPE_MISC, SC_PIPE_FRAMES,
    SC_PRE_COEF, SC_PRE_TMP,                              // PIL-bicubic letterbox: coefficient tables, uint8 intermediate
    SC_PIPE_LB                                            // chained det -> pose pipeline: the detector's letterboxed input tensor
};

void hbp_set_error(const char* fmt, ...);
int hbp_cuda_fail(cudaError_t e, const char* what, const char* file, int line);
void* hbp_scratch(hbp_ctx* ctx, int slot, size_t bytes);      // nullptr on failure (error set)
void* hbp_pinned(hbp_ctx* ctx, size_t bytes);

#define HBP_CUDA(expr)                                                         \
    do {                                                                       \
        cudaError_t e__ = (expr);                                              \
        if (e__ != cudaSuccess) return hbp_cuda_fail(e__, #expr, __FILE__, __LINE__); \
    } while (0)

#define HBP_REQUIRE(cond, msg)                                                 \
    do {                                                                       \
        if (!(cond)) {                                                         \
            hbp_set_error("%s: %s", __func__, msg);                            \
            return HBP_ERR_INVALID;                                            \
        }                                                                      \
    } while (0)

#define HBP_LAUNCH_CHECK(ctx)                                                  \
    do {                                                                       \
        (ctx)->launches++;                                                     \
        cudaError_t e__ = cudaGetLastError();                                  \
        if (e__ != cudaSuccess) return hbp_cuda_fail(e__, "kernel launch", __FILE__, __LINE__); \
    } while (0)

// Helper for host-mode calls: copies a host array into a scratch slot and
// returns the device pointer (or the pointer itself in device mode).
struct Stager {
    hbp_ctx* ctx;
    int mem;
    int status = HBP_OK;
    Stager(hbp_ctx* c, int m) : ctx(c), mem(m) {}
    template <typename T>
    const T* in(const T* p, size_t count, int slot) {
        if (mem == HBP_DEVICE || p == nullptr || count == 0) return p;
        void* d = hbp_scratch(ctx, slot, count * sizeof(T));
        if (!d) { status = HBP_ERR_NOMEM; return nullptr; }
        cudaError_t e = cudaMemcpyAsync(d, p, count * sizeof(T), cudaMemcpyHostToDevice, ctx->stream);
        if (e != cudaSuccess) status = hbp_cuda_fail(e, "stage-in", __FILE__, __LINE__);
        return static_cast<const T*>(d);
    }
    template <typename T>
    T* out(T* p, size_t count, int slot) {
        if (mem == HBP_DEVICE || p == nullptr || count == 0) return p;
        void* d = hbp_scratch(ctx, slot, count * sizeof(T));
        if (!d) { status = HBP_ERR_NOMEM; return nullptr; }
        return static_cast<T*>(d);
    }
    template <typename T>
    void back(T* host, const T* dev, size_t count) {
        if (mem == HBP_DEVICE || host == nullptr || count == 0 || status != HBP_OK) return;
        cudaError_t e = cudaMemcpyAsync(host, dev, count * sizeof(T), cudaMemcpyDeviceToHost, ctx->stream);
        if (e != cudaSuccess) status = hbp_cuda_fail(e, "stage-out", __FILE__, __LINE__);
    }
    int finish() {
        if (status != HBP_OK) return status;
        if (mem == HBP_HOST) {
            cudaError_t e = cudaStreamSynchronize(ctx->stream);
            if (e != cudaSuccess) return hbp_cuda_fail(e, "sync", __FILE__, __LINE__);
        }
        return HBP_OK;
    }
};

// ---- stage launchers (device pointers only; implemented per .cu) -----------
int k_preprocess(hbp_ctx*, const uint8_t* frames, int n, int h, int w, int mode, int out_h,
                 int out_w, int swap_rb, int pad_value, void* out, int out_dtype, int out_layout);
int k_yolo_decode_raw(hbp_ctx*, const float* h0, const float* h1, const float* h2, int B, int s0,
                      int s1, int s2, int nc, int in_w, int in_h, float* out);
int k_yolo_nms(hbp_ctx*, const float* pred, int B, int N, int nc, float conf, double iou,
               const int* classes, int n_classes, int max_det, float* out_det, int* out_count);
int k_yolo_nms_legacy(hbp_ctx*, const float* pred, int B, int N, int nc, float conf, float thr,
                      int max_out, float* out_det, int* out_count);
int k_scale_coords(hbp_ctx*, float* boxes, int n, int h1, int w1, int h0, int w0);
int k_edet_filter(hbp_ctx*, const float* boxes, const float* scores, const float* classes, int F,
                  int K, float person_class, float thr, float xe, float ye, int img_h, int img_w,
                  int max_persons, float* out_boxes, int* out_count);
int k_crop_warp(hbp_ctx*, const uint8_t* frames, int n_frames, int h, int w, const double* M,
                const int* frame_idx, int P, int out_h, int out_w, int swap_rb, void* out,
                int out_dtype, const int* live = nullptr);   // live: optional device count, slots >= *live are skipped
int k_keypoint_lengths(hbp_ctx*, const float* kpts, const uint32_t* ignored, const double* pixel_to_cm, int P,
                       float* lengths, double* torso);
int k_decode_proportions(hbp_ctx*, const void* hm, int dtype, int P, int J, int Hh, int Wh,
                         const float* boxes, const double* height_cm, const float* thr,
                         int quarter, float* kpts_hm, float* kpts_img, float* scores,
                         int32_t* idx, uint32_t* ignored, float* lengths, double* torso,
                         const int* live = nullptr,          // optional device count, slots >= *live are skipped
                         const double* Maff = nullptr, int crop_h = 0, int crop_w = 0);   // optional (P,6) dst->src matrices: general inverse-affine remap
int k_yolo_filter(hbp_ctx*, const float* pred, int B, int N, int nc, float conf, const int* classes, int n_classes,
                  int cand_cap, int* out_count);
int k_yolo_nms_bounded(hbp_ctx*, const float* pred, int B, int N, int nc, float conf, double iou,
                       const int* classes, int n_classes, int max_det, int cand_cap, float* out_det, int* out_count,
                       int* status);
int k_persons_from_yolo(hbp_ctx*, const float* det, const int* det_count, int F, int max_det, int in_h, int in_w,
                        int img_h, int img_w, int out_h, int out_w, const double* heights, int n_heights,
                        int persons_cap, double* M, float* boxes, int* frame_idx, double* height_cm,
                        int* n_persons, int* status);
int k_persons_from_edet(hbp_ctx*, const float* boxes_n, const int* counts, int F, int max_persons, int img_h,
                        int img_w, int out_h, int out_w, const double* heights, int n_heights, int persons_cap,
                        double* M, float* boxes, int* frame_idx, double* height_cm, int* n_persons, int* status);
// hrnet.cu
int hrnet_load(hbp_ctx*, int width, int in_h, int in_w, const void* w16, size_t nw,
               const float* bias, size_t nb);
int hrnet_forward(hbp_ctx*, const __half* crops_dev, int P, void* heatmaps_dev, int out_dtype);
int hrnet_set_engine(hbp_ctx*, int engine);
int hrnet_debug_tensor(hbp_ctx*, int id, void* out_host, size_t max_bytes, int* n, int* h, int* w, int* c);
int hrnet_forward_until(hbp_ctx*, const __half* crops_dev, int P, int stop_after);
int hrnet_op_count(hbp_ctx*);
const char* hrnet_op_name(hbp_ctx*, int id);
void hrnet_free(hbp_ctx*);
void hrnet_dims(hbp_ctx*, int* in_h, int* in_w, int* width);
int hrnet_single_conv(hbp_ctx* ctx, int engine, const __half* in, int P, int H, int W, int Cin,
                      const __half* w, const float* bias, const __half* res, int Cout, int k, int stride,
                      int up, int relu, __half* out, int* used_engine, int iters = 0, float* avg_ms = nullptr);
