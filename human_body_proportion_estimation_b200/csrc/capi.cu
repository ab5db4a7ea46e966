// extern "C" boundary of libhbp_b200.so (include/hbp.h): context, memory,
// timing, and the host/device staging around every stage launcher.
#include "hbp_internal.cuh"
#include <cstdarg>
#include <cstdio>
#include <cstring>

static thread_local char g_err[1024] = "";

void hbp_set_error(const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
}

int hbp_cuda_fail(cudaError_t e, const char* what, const char* file, int line) {
    hbp_set_error("CUDA error %d (%s) at %s:%d in `%s`", (int)e, cudaGetErrorString(e), file, line, what);
    return e == cudaErrorMemoryAllocation ? HBP_ERR_NOMEM : HBP_ERR_CUDA;
}

void* hbp_scratch(hbp_ctx* ctx, int slot, size_t bytes) {
    if (bytes == 0) bytes = 256;
    if (ctx->scratch_bytes[slot] >= bytes) return ctx->scratch[slot];
    if (ctx->scratch[slot]) {
        cudaStreamSynchronize(ctx->stream);
        cudaFree(ctx->scratch[slot]);
        ctx->scratch[slot] = nullptr;
        ctx->scratch_bytes[slot] = 0;
    }
    size_t want = (bytes + (bytes >> 2) + 4095) & ~size_t(4095);   // 25 % headroom
    void* p = nullptr;
    cudaError_t e = cudaMalloc(&p, want);
    if (e != cudaSuccess) {
        hbp_cuda_fail(e, "cudaMalloc(scratch)", __FILE__, __LINE__);
        return nullptr;
    }
    ctx->scratch[slot] = p;
    ctx->scratch_bytes[slot] = want;
    return p;
}

void* hbp_pinned(hbp_ctx* ctx, size_t bytes) {
    if (ctx->pinned_bytes >= bytes) return ctx->pinned;
    if (ctx->pinned) { cudaStreamSynchronize(ctx->stream); cudaFreeHost(ctx->pinned); ctx->pinned = nullptr; ctx->pinned_bytes = 0; }
    size_t want = (bytes * 2 + 4095) & ~size_t(4095);
    void* p = nullptr;
    cudaError_t e = cudaMallocHost(&p, want);
    if (e != cudaSuccess) { hbp_cuda_fail(e, "cudaMallocHost", __FILE__, __LINE__); return nullptr; }
    ctx->pinned = p;
    ctx->pinned_bytes = want;
    return p;
}

static int bind(hbp_ctx* ctx) {
    if (!ctx) { hbp_set_error("null context"); return HBP_ERR_INVALID; }
    HBP_CUDA(cudaSetDevice(ctx->device));
    return HBP_OK;
}
#define BIND(ctx) do { int s__ = bind(ctx); if (s__ != HBP_OK) return s__; } while (0)

extern "C" {

int hbp_version(void) { return HBP_VERSION; }
const char* hbp_last_error(void) { return g_err; }

int hbp_device_count(int* n) {
    if (!n) return HBP_ERR_INVALID;
    cudaError_t e = cudaGetDeviceCount(n);
    if (e != cudaSuccess) { *n = 0; return hbp_cuda_fail(e, "cudaGetDeviceCount", __FILE__, __LINE__); }
    return HBP_OK;
}

int hbp_ctx_create(int device, hbp_ctx** out) {
    if (!out) return HBP_ERR_INVALID;
    *out = nullptr;
    int n = 0;
    cudaError_t e = cudaGetDeviceCount(&n);
    if (e != cudaSuccess || n == 0) {
        hbp_set_error("no CUDA device available (%s); this library has no CPU fallback",
                      e == cudaSuccess ? "count == 0" : cudaGetErrorString(e));
        return HBP_ERR_CUDA;
    }
    if (device < 0 || device >= n) { hbp_set_error("device %d out of range (0..%d)", device, n - 1); return HBP_ERR_INVALID; }
    HBP_CUDA(cudaSetDevice(device));
    cudaDeviceProp prop;
    HBP_CUDA(cudaGetDeviceProperties(&prop, device));
    if (prop.major != 10) {
        hbp_set_error("device %d is sm_%d%d; this library is built for sm_100a (B200) only", device, prop.major, prop.minor);
        return HBP_ERR_STATE;
    }
    hbp_ctx* ctx = new hbp_ctx();
    ctx->device = device;
    ctx->sm_count = prop.multiProcessorCount;
    HBP_CUDA(cudaStreamCreateWithFlags(&ctx->stream, cudaStreamNonBlocking));
    for (int i = 0; i < HBP_TIMER_SLOTS; ++i) {
        HBP_CUDA(cudaEventCreate(&ctx->ev_start[i]));
        HBP_CUDA(cudaEventCreate(&ctx->ev_stop[i]));
    }
    *out = ctx;
    return HBP_OK;
}

int hbp_ctx_destroy(hbp_ctx* ctx) {
    if (!ctx) return HBP_OK;
    cudaSetDevice(ctx->device);
    cudaStreamSynchronize(ctx->stream);
    hrnet_free(ctx);
    for (int i = 0; i < HBP_SCRATCH_SLOTS; ++i) if (ctx->scratch[i]) cudaFree(ctx->scratch[i]);
    if (ctx->pinned) cudaFreeHost(ctx->pinned);
    if (ctx->l2_flush) cudaFree(ctx->l2_flush);
    for (hbp_pipe_slot& sl : ctx->pipe) {
        if (sl.d_frames) cudaFree(sl.d_frames);
        if (sl.d_misc) cudaFree(sl.d_misc);
        if (sl.h_pin) cudaFreeHost(sl.h_pin);
        if (sl.ev_h2d) cudaEventDestroy(sl.ev_h2d);
        if (sl.ev_crop) cudaEventDestroy(sl.ev_crop);
        if (sl.ev_done) cudaEventDestroy(sl.ev_done);
    }
    if (ctx->copy_stream) cudaStreamDestroy(ctx->copy_stream);
    for (int i = 0; i < HBP_TIMER_SLOTS; ++i) { cudaEventDestroy(ctx->ev_start[i]); cudaEventDestroy(ctx->ev_stop[i]); }
    cudaStreamDestroy(ctx->stream);
    delete ctx;
    return HBP_OK;
}

int hbp_sync(hbp_ctx* ctx) { BIND(ctx); HBP_CUDA(cudaStreamSynchronize(ctx->stream)); return HBP_OK; }

int hbp_dev_alloc(hbp_ctx* ctx, size_t nbytes, void** out) {
    BIND(ctx);
    if (!out) return HBP_ERR_INVALID;
    HBP_CUDA(cudaMalloc(out, nbytes ? nbytes : 256));
    return HBP_OK;
}
int hbp_dev_free(hbp_ctx* ctx, void* p) { BIND(ctx); HBP_CUDA(cudaStreamSynchronize(ctx->stream)); HBP_CUDA(cudaFree(p)); return HBP_OK; }
int hbp_host_alloc(hbp_ctx* ctx, size_t nbytes, void** out) {
    BIND(ctx);
    if (!out) return HBP_ERR_INVALID;
    HBP_CUDA(cudaMallocHost(out, nbytes ? nbytes : 256));
    return HBP_OK;
}
int hbp_host_free(hbp_ctx* ctx, void* p) { BIND(ctx); HBP_CUDA(cudaFreeHost(p)); return HBP_OK; }
int hbp_copy_h2d(hbp_ctx* ctx, void* d, const void* s, size_t n) {
    BIND(ctx); HBP_CUDA(cudaMemcpyAsync(d, s, n, cudaMemcpyHostToDevice, ctx->stream)); return HBP_OK;
}
int hbp_copy_d2h(hbp_ctx* ctx, void* d, const void* s, size_t n) {
    BIND(ctx); HBP_CUDA(cudaMemcpyAsync(d, s, n, cudaMemcpyDeviceToHost, ctx->stream)); return HBP_OK;
}
int hbp_memset_dev(hbp_ctx* ctx, void* d, int byte, size_t n) {
    BIND(ctx); HBP_CUDA(cudaMemsetAsync(d, byte, n, ctx->stream)); return HBP_OK;
}

int hbp_timer_start(hbp_ctx* ctx, int slot) {
    BIND(ctx);
    HBP_REQUIRE(slot >= 0 && slot < HBP_TIMER_SLOTS, "bad timer slot");
    HBP_CUDA(cudaEventRecord(ctx->ev_start[slot], ctx->stream));
    return HBP_OK;
}
int hbp_timer_stop(hbp_ctx* ctx, int slot) {
    BIND(ctx);
    HBP_REQUIRE(slot >= 0 && slot < HBP_TIMER_SLOTS, "bad timer slot");
    HBP_CUDA(cudaEventRecord(ctx->ev_stop[slot], ctx->stream));
    return HBP_OK;
}
int hbp_timer_elapsed_ms(hbp_ctx* ctx, int slot, float* ms) {
    BIND(ctx);
    HBP_REQUIRE(slot >= 0 && slot < HBP_TIMER_SLOTS && ms, "bad timer slot");
    HBP_CUDA(cudaEventSynchronize(ctx->ev_stop[slot]));
    HBP_CUDA(cudaEventElapsedTime(ms, ctx->ev_start[slot], ctx->ev_stop[slot]));
    return HBP_OK;
}
int hbp_flush_l2(hbp_ctx* ctx) {
    BIND(ctx);
    const size_t n = size_t(256) << 20;
    if (!ctx->l2_flush) HBP_CUDA(cudaMalloc(&ctx->l2_flush, n));
    HBP_CUDA(cudaMemsetAsync(ctx->l2_flush, 1, n, ctx->stream));
    return HBP_OK;
}
int hbp_kernel_launches(hbp_ctx* ctx, uint64_t* n) {
    if (!ctx || !n) return HBP_ERR_INVALID;
    *n = ctx->launches;
    return HBP_OK;
}

// ---------------------------------------------------------------------------
int hbp_preprocess(hbp_ctx* ctx, const uint8_t* frames, int n, int h, int w, int mode, int out_h,
                   int out_w, int swap_rb, int pad_value, void* out, int out_dtype, int out_layout, int mem) {
    BIND(ctx);
    HBP_REQUIRE(frames && out && n > 0 && h > 0 && w > 0 && out_h > 0 && out_w > 0, "bad shape");
    HBP_REQUIRE(mode >= HBP_PRE_COPY && mode <= HBP_PRE_LETTERBOX_PIL, "bad mode");
    HBP_REQUIRE(out_dtype >= HBP_U8 && out_dtype <= HBP_F32, "bad dtype");
    HBP_REQUIRE(mode != HBP_PRE_COPY || (out_h == h && out_w == w), "COPY mode needs out size == in size");
    size_t esz = out_dtype == HBP_U8 ? 1 : out_dtype == HBP_F16 ? 2 : 4;
    size_t in_n = (size_t)n * h * w * 3, out_n = (size_t)n * out_h * out_w * 3 * esz;
    Stager st(ctx, mem);
    const uint8_t* d_in = st.in(frames, in_n, SC_IN0);
    uint8_t* d_out = st.out((uint8_t*)out, out_n, SC_OUT0);
    if (st.status) return st.status;
    int s = k_preprocess(ctx, d_in, n, h, w, mode, out_h, out_w, swap_rb, pad_value, d_out, out_dtype, out_layout);
    if (s) return s;
    st.back((uint8_t*)out, d_out, out_n);
    return st.finish();
}

int hbp_yolo_decode_raw(hbp_ctx* ctx, const float* h0, const float* h1, const float* h2, int B,
                        int s0, int s1, int s2, int nc, int in_w, int in_h, float* out, int mem) {
    BIND(ctx);
    HBP_REQUIRE(h0 && h1 && h2 && out && B > 0 && nc > 0 && s0 > 0 && s1 > 0 && s2 > 0, "bad shape");
    size_t E = 5 + nc;
    size_t n0 = (size_t)B * 3 * s0 * s0 * E, n1 = (size_t)B * 3 * s1 * s1 * E, n2 = (size_t)B * 3 * s2 * s2 * E;
    Stager st(ctx, mem);
    const float* d0 = st.in(h0, n0, SC_IN0);
    const float* d1 = st.in(h1, n1, SC_IN1);
    const float* d2 = st.in(h2, n2, SC_IN2);
    float* d_out = st.out(out, n0 + n1 + n2, SC_OUT0);
    if (st.status) return st.status;
    int s = k_yolo_decode_raw(ctx, d0, d1, d2, B, s0, s1, s2, nc, in_w, in_h, d_out);
    if (s) return s;
    st.back(out, d_out, n0 + n1 + n2);
    return st.finish();
}

int hbp_yolo_nms(hbp_ctx* ctx, const float* pred, int B, int N, int nc, float conf, double iou,
                 const int* classes, int n_classes, int max_det, float* out_det, int* out_count, int mem) {
    BIND(ctx);
    HBP_REQUIRE(pred && out_det && out_count && B > 0 && N > 0 && nc > 0 && max_det > 0, "bad shape");
    HBP_REQUIRE(conf >= 0.f && conf <= 1.f, "Invalid Confidence threshold, valid values are between 0.0 and 1.0");
    HBP_REQUIRE(iou >= 0.0 && iou <= 1.0, "Invalid IoU, valid values are between 0.0 and 1.0");
    HBP_REQUIRE(n_classes >= 0 && (n_classes == 0 || classes), "bad class filter");
    Stager st(ctx, mem);
    const float* d_pred = st.in(pred, (size_t)B * N * (5 + nc), SC_IN0);
    const int* d_cls = st.in(classes, (size_t)n_classes, SC_IN1);
    float* d_det = st.out(out_det, (size_t)B * max_det * 6, SC_OUT0);
    int* d_cnt = st.out(out_count, (size_t)B, SC_OUT1);
    if (st.status) return st.status;
    int s = k_yolo_nms(ctx, d_pred, B, N, nc, conf, iou, d_cls, n_classes, max_det, d_det, d_cnt);
    if (s) return s;
    st.back(out_det, d_det, (size_t)B * max_det * 6);
    st.back(out_count, d_cnt, (size_t)B);
    return st.finish();
}

int hbp_yolo_filter(hbp_ctx* ctx, const float* pred, int B, int N, int nc, float conf, const int* classes, int n_classes,
                    int cand_cap, int* out_count, int mem) {
    BIND(ctx);
    HBP_REQUIRE(pred && out_count && B > 0 && N > 0 && nc > 0 && cand_cap > 0, "bad shape");
    HBP_REQUIRE(n_classes >= 0 && (n_classes == 0 || classes), "bad class filter");
    Stager st(ctx, mem);
    const float* d_pred = st.in(pred, (size_t)B * N * (5 + nc), SC_IN0);
    const int* d_cls = st.in(classes, (size_t)n_classes, SC_IN1);
    int* d_cnt = st.out(out_count, (size_t)B, SC_OUT1);
    if (st.status) return st.status;
    int s = k_yolo_filter(ctx, d_pred, B, N, nc, conf, d_cls, n_classes, cand_cap, d_cnt);
    if (s) return s;
    st.back(out_count, d_cnt, (size_t)B);
    return st.finish();
}

int hbp_yolo_nms_legacy(hbp_ctx* ctx, const float* pred, int B, int N, int nc, float conf, float thr,
                        int max_out, float* out_det, int* out_count, int mem) {
    BIND(ctx);
    HBP_REQUIRE(pred && out_det && out_count && B > 0 && N > 0 && nc > 0 && max_out > 0, "bad shape");
    Stager st(ctx, mem);
    const float* d_pred = st.in(pred, (size_t)B * N * (5 + nc), SC_IN0);
    float* d_det = st.out(out_det, (size_t)B * max_out * 7, SC_OUT0);
    int* d_cnt = st.out(out_count, (size_t)B, SC_OUT1);
    if (st.status) return st.status;
    int s = k_yolo_nms_legacy(ctx, d_pred, B, N, nc, conf, thr, max_out, d_det, d_cnt);
    if (s) return s;
    st.back(out_det, d_det, (size_t)B * max_out * 7);
    st.back(out_count, d_cnt, (size_t)B);
    return st.finish();
}

int hbp_scale_coords(hbp_ctx* ctx, float* boxes, int n, int h1, int w1, int h0, int w0, int mem) {
    BIND(ctx);
    HBP_REQUIRE(n >= 0 && h1 > 0 && w1 > 0 && h0 > 0 && w0 > 0, "bad shape");
    if (n == 0) return HBP_OK;
    HBP_REQUIRE(boxes, "null boxes");
    Stager st(ctx, mem);
    float* d = const_cast<float*>(st.in((const float*)boxes, (size_t)n * 4, SC_IN0));
    if (st.status) return st.status;
    int s = k_scale_coords(ctx, d, n, h1, w1, h0, w0);
    if (s) return s;
    st.back(boxes, d, (size_t)n * 4);
    return st.finish();
}

int hbp_edet_person_filter(hbp_ctx* ctx, const float* boxes, const float* scores, const float* classes,
                           int F, int K, float person_class, float thr, float xe, float ye, int img_h,
                           int img_w, int max_persons, float* out_boxes, int* out_count, int mem) {
    BIND(ctx);
    HBP_REQUIRE(boxes && scores && classes && out_boxes && out_count && F > 0 && K > 0 && max_persons > 0, "bad shape");
    Stager st(ctx, mem);
    const float* db = st.in(boxes, (size_t)F * K * 4, SC_IN0);
    const float* ds = st.in(scores, (size_t)F * K, SC_IN1);
    const float* dc = st.in(classes, (size_t)F * K, SC_IN2);
    float* dob = st.out(out_boxes, (size_t)F * max_persons * 4, SC_OUT0);
    int* doc = st.out(out_count, (size_t)F, SC_OUT1);
    if (st.status) return st.status;
    int s = k_edet_filter(ctx, db, ds, dc, F, K, person_class, thr, xe, ye, img_h, img_w, max_persons, dob, doc);
    if (s) return s;
    st.back(out_boxes, dob, (size_t)F * max_persons * 4);
    st.back(out_count, doc, (size_t)F);
    return st.finish();
}

int hbp_crop_warp(hbp_ctx* ctx, const uint8_t* frames, int n_frames, int h, int w, const double* M,
                  const int* frame_idx, int P, int out_h, int out_w, int swap_rb, void* out,
                  int out_dtype, int mem) {
    BIND(ctx);
    HBP_REQUIRE(frames && M && frame_idx && out && n_frames > 0 && h > 0 && w > 0 && P >= 0 && out_h > 0 && out_w > 0, "bad shape");
    HBP_REQUIRE(out_dtype == HBP_F16 || out_dtype == HBP_F32, "crop output must be f16 or f32");
    if (P == 0) return HBP_OK;
    if (mem == HBP_HOST) {
        for (int p = 0; p < P; ++p)
            HBP_REQUIRE(frame_idx[p] >= 0 && frame_idx[p] < n_frames, "frame_idx out of range");
    }
    size_t esz = out_dtype == HBP_F16 ? 2 : 4;
    size_t out_n = (size_t)P * 3 * out_h * out_w * esz;
    Stager st(ctx, mem);
    const uint8_t* df = st.in(frames, (size_t)n_frames * h * w * 3, SC_IN0);
    const double* dM = st.in(M, (size_t)P * 6, SC_IN1);
    const int* dfi = st.in(frame_idx, (size_t)P, SC_IN2);
    uint8_t* dout = st.out((uint8_t*)out, out_n, SC_OUT0);
    if (st.status) return st.status;
    int s = k_crop_warp(ctx, df, n_frames, h, w, dM, dfi, P, out_h, out_w, swap_rb, dout, out_dtype);
    if (s) return s;
    st.back((uint8_t*)out, dout, out_n);
    return st.finish();
}

int hbp_hrnet_load(hbp_ctx* ctx, int width, int in_h, int in_w, const void* w16, size_t nw,
                   const float* bias, size_t nb) {
    BIND(ctx);
    HBP_REQUIRE(w16 && bias, "null weights");
    HBP_REQUIRE(width == 32 || width == 48, "HRNet width must be 32 or 48");
    HBP_REQUIRE(in_h > 0 && in_w > 0 && in_h % 32 == 0 && in_w % 32 == 0, "input size must be a multiple of 32");
    return hrnet_load(ctx, width, in_h, in_w, w16, nw, bias, nb);
}

static int conv2d_impl(hbp_ctx* ctx, int engine, const void* in, int P, int H, int W, int Cin, const void* weights,
                       const float* bias, const void* residual, int Cout, int k, int stride, int up, int relu,
                       void* out, int* used_engine, int mem, int iters, float* avg_ms);

int hbp_conv2d_nhwc(hbp_ctx* ctx, int engine, const void* in, int P, int H, int W, int Cin, const void* weights,
                    const float* bias, const void* residual, int Cout, int k, int stride, int up, int relu,
                    void* out, int* used_engine, int mem) {
    return conv2d_impl(ctx, engine, in, P, H, W, Cin, weights, bias, residual, Cout, k, stride, up, relu, out,
                       used_engine, mem, 0, nullptr);
}

int hbp_conv2d_nhwc_timed(hbp_ctx* ctx, int engine, const void* in, int P, int H, int W, int Cin, const void* weights,
                          const float* bias, const void* residual, int Cout, int k, int stride, int up, int relu,
                          void* out, int* used_engine, int mem, int iters, float* avg_ms) {
    HBP_REQUIRE(iters > 0 && avg_ms, "iters > 0 and avg_ms required");
    return conv2d_impl(ctx, engine, in, P, H, W, Cin, weights, bias, residual, Cout, k, stride, up, relu, out,
                       used_engine, mem, iters, avg_ms);
}

}  // extern "C"

static int conv2d_impl(hbp_ctx* ctx, int engine, const void* in, int P, int H, int W, int Cin, const void* weights,
                       const float* bias, const void* residual, int Cout, int k, int stride, int up, int relu,
                       void* out, int* used_engine, int mem, int iters, float* avg_ms) {
    BIND(ctx);
    HBP_REQUIRE(in && weights && bias && out && P > 0 && H > 0 && W > 0 && Cin > 0 && Cout > 0, "bad shape");
    HBP_REQUIRE((k == 1 || k == 3) && (stride == 1 || stride == 2) && up >= 1 && Cin % 8 == 0 && Cout % 4 == 0, "unsupported conv");
    HBP_REQUIRE(H % stride == 0 && W % stride == 0, "size not divisible by stride");
    const size_t n_in = (size_t)P * H * W * Cin, n_w = (size_t)k * k * Cout * Cin;
    const size_t n_out = (size_t)P * (H / stride * up) * (W / stride * up) * Cout;
    Stager st(ctx, mem);
    const __half* d_in = st.in((const __half*)in, n_in, SC_IN0);
    const __half* d_w = st.in((const __half*)weights, n_w, SC_IN1);
    const float* d_b = st.in(bias, (size_t)Cout, SC_IN2);
    const __half* d_r = st.in((const __half*)residual, n_out, SC_IN3);
    __half* d_o = st.out((__half*)out, n_out, SC_OUT0);
    if (st.status) return st.status;
    int s = hrnet_single_conv(ctx, engine, d_in, P, H, W, Cin, d_w, d_b, d_r, Cout, k, stride, up, relu, d_o, used_engine,
                              iters, avg_ms);
    if (s) return s;
    st.back((__half*)out, d_o, n_out);
    return st.finish();
}

extern "C" {

int hbp_hrnet_set_engine(hbp_ctx* ctx, int engine) { BIND(ctx); return hrnet_set_engine(ctx, engine); }

int hbp_hrnet_debug_tensor(hbp_ctx* ctx, int id, void* out_host, size_t max_bytes, int* n, int* h, int* w, int* c) {
    BIND(ctx);
    return hrnet_debug_tensor(ctx, id, out_host, max_bytes, n, h, w, c);
}

int hbp_hrnet_forward_until(hbp_ctx* ctx, const void* crops, int P, int op_index, int mem) {
    BIND(ctx);
    HBP_REQUIRE(crops && P > 0, "bad arguments");
    if (!ctx->hrnet) { hbp_set_error("hbp_hrnet_forward_until before hbp_hrnet_load"); return HBP_ERR_STATE; }
    int ih, iw, wd;
    hrnet_dims(ctx, &ih, &iw, &wd);
    Stager st(ctx, mem);
    const __half* din = st.in((const __half*)crops, (size_t)P * 3 * ih * iw, SC_IN0);
    if (st.status) return st.status;
    return hrnet_forward_until(ctx, din, P, op_index);
}

int hbp_hrnet_op_name(hbp_ctx* ctx, int op_index, char* buf, size_t buf_bytes, int* n_ops) {
    BIND(ctx);
    if (!ctx->hrnet) { hbp_set_error("hbp_hrnet_op_name before hbp_hrnet_load"); return HBP_ERR_STATE; }
    if (n_ops) *n_ops = hrnet_op_count(ctx);
    if (!buf || buf_bytes == 0) return HBP_OK;
    const char* nm = hrnet_op_name(ctx, op_index);
    HBP_REQUIRE(nm != nullptr, "op index out of range");
    snprintf(buf, buf_bytes, "%s", nm);
    return HBP_OK;
}

static int decode_impl(hbp_ctx* ctx, const void* hm, int dtype, int P, int J, int Hh, int Wh,
                       const float* boxes, const double* M, int crop_h, int crop_w, const double* height_cm, const float* thr, int quarter,
                       float* kpts_hm, float* kpts_img, float* scores, int32_t* idx,
                       uint32_t* ignored, float* lengths, double* torso, int mem);

int hbp_decode_proportions(hbp_ctx* ctx, const void* hm, int dtype, int P, int J, int Hh, int Wh,
                           const float* boxes, const double* height_cm, const float* thr, int quarter,
                           float* kpts_hm, float* kpts_img, float* scores, int32_t* idx,
                           uint32_t* ignored, float* lengths, double* torso, int mem) {
    return decode_impl(ctx, hm, dtype, P, J, Hh, Wh, boxes, nullptr, 0, 0, height_cm, thr, quarter, kpts_hm, kpts_img, scores, idx,
                       ignored, lengths, torso, mem);
}

int hbp_decode_proportions_affine(hbp_ctx* ctx, const void* hm, int dtype, int P, int J, int Hh, int Wh,
                                  const float* boxes, const double* M, int crop_h, int crop_w, const double* height_cm,
                                  const float* thr, int quarter, float* kpts_hm, float* kpts_img, float* scores, int32_t* idx,
                                  uint32_t* ignored, float* lengths, double* torso, int mem) {
    if (!M || !boxes || crop_h <= 0 || crop_w <= 0) { hbp_set_error("hbp_decode_proportions_affine: M, boxes and the crop size are required"); return HBP_ERR_INVALID; }
    return decode_impl(ctx, hm, dtype, P, J, Hh, Wh, boxes, M, crop_h, crop_w, height_cm, thr, quarter, kpts_hm, kpts_img, scores, idx,
                       ignored, lengths, torso, mem);
}

static int decode_impl(hbp_ctx* ctx, const void* hm, int dtype, int P, int J, int Hh, int Wh,
                       const float* boxes, const double* M, int crop_h, int crop_w, const double* height_cm, const float* thr, int quarter,
                       float* kpts_hm, float* kpts_img, float* scores, int32_t* idx,
                       uint32_t* ignored, float* lengths, double* torso, int mem) {
    BIND(ctx);
    HBP_REQUIRE(hm && P >= 0 && J > 0 && J <= 32 && Hh > 0 && Wh > 0, "bad shape");
    HBP_REQUIRE(dtype == HBP_F32 || dtype == HBP_F16, "heatmaps must be f32 or f16");
    HBP_REQUIRE((size_t)Hh * Wh < (size_t(1) << 24), "heatmap too large for exact float32 indices");
    HBP_REQUIRE(!boxes || (height_cm && thr), "boxes need height_cm and joint_thr");
    HBP_REQUIRE(!(lengths || torso) || (boxes && J == 17), "lengths need boxes and J == 17");
    HBP_REQUIRE(!(kpts_img || ignored) || boxes, "image-space outputs need boxes");
    if (P == 0) return HBP_OK;
    size_t esz = dtype == HBP_F32 ? 4 : 2;
    Stager st(ctx, mem);
    const uint8_t* dhm = st.in((const uint8_t*)hm, (size_t)P * J * Hh * Wh * esz, SC_IN0);
    const float* db = st.in(boxes, (size_t)P * 4, SC_IN1);
    const double* dh = st.in(height_cm, (size_t)P, SC_IN2);
    const float* dt = st.in(thr, (size_t)J, SC_IN3);
    const double* dM = st.in(M, (size_t)P * 6, SC_IN4);
    float* o_hm = st.out(kpts_hm, (size_t)P * J * 2, SC_OUT0);
    float* o_img = st.out(kpts_img, (size_t)P * J * 2, SC_OUT1);
    float* o_sc = st.out(scores, (size_t)P * J, SC_OUT2);
    int32_t* o_idx = st.out(idx, (size_t)P * J, SC_OUT3);
    uint32_t* o_ig = st.out(ignored, (size_t)P, SC_OUT4);
    float* o_len = st.out(lengths, (size_t)P * 11, SC_OUT5);
    double* o_to = st.out(torso, (size_t)P, SC_OUT6);
    if (st.status) return st.status;
    int s = k_decode_proportions(ctx, dhm, dtype, P, J, Hh, Wh, db, dh, dt, quarter, o_hm, o_img, o_sc,
                                 o_idx, o_ig, o_len, o_to, nullptr, dM, crop_h, crop_w);
    if (s) return s;
    st.back(kpts_hm, o_hm, (size_t)P * J * 2);
    st.back(kpts_img, o_img, (size_t)P * J * 2);
    st.back(scores, o_sc, (size_t)P * J);
    st.back(idx, o_idx, (size_t)P * J);
    st.back(ignored, o_ig, (size_t)P);
    st.back(lengths, o_len, (size_t)P * 11);
    st.back(torso, o_to, (size_t)P);
    return st.finish();
}

int hbp_keypoint_lengths(hbp_ctx* ctx, const float* kpts_img, const uint32_t* ignored, const double* pixel_to_cm, int P,
                         float* lengths, double* torso, int mem) {
    BIND(ctx);
    HBP_REQUIRE(kpts_img && pixel_to_cm && P >= 0 && (lengths || torso), "bad arguments");
    if (P == 0) return HBP_OK;
    Stager st(ctx, mem);
    const float* dk = st.in(kpts_img, (size_t)P * 34, SC_IN0);
    const uint32_t* di = st.in(ignored, (size_t)P, SC_IN1);
    const double* dp = st.in(pixel_to_cm, (size_t)P, SC_IN2);
    float* o_len = st.out(lengths, (size_t)P * 11, SC_OUT0);
    double* o_to = st.out(torso, (size_t)P, SC_OUT1);
    if (st.status) return st.status;
    int s = k_keypoint_lengths(ctx, dk, di, dp, P, o_len, o_to);
    if (s) return s;
    st.back(lengths, o_len, (size_t)P * 11);
    st.back(torso, o_to, (size_t)P);
    return st.finish();
}

int hbp_hrnet_forward(hbp_ctx* ctx, const void* crops, int P, void* heatmaps, int out_dtype, int mem) {
    BIND(ctx);
    HBP_REQUIRE(crops && heatmaps && P > 0, "bad arguments");
    HBP_REQUIRE(out_dtype == HBP_F16 || out_dtype == HBP_F32, "heatmaps must be f16 or f32");
    if (!ctx->hrnet) { hbp_set_error("hbp_hrnet_forward before hbp_hrnet_load"); return HBP_ERR_STATE; }
    int ih, iw, wd;
    hrnet_dims(ctx, &ih, &iw, &wd);
    size_t in_n = (size_t)P * 3 * ih * iw;
    size_t out_n = (size_t)P * 17 * (ih / 4) * (iw / 4) * (out_dtype == HBP_F16 ? 2 : 4);
    Stager st(ctx, mem);
    const __half* din = st.in((const __half*)crops, in_n, SC_IN0);
    uint8_t* dout = st.out((uint8_t*)heatmaps, out_n, SC_OUT0);
    if (st.status) return st.status;
    int s = hrnet_forward(ctx, din, P, dout, out_dtype);
    if (s) return s;
    st.back((uint8_t*)heatmaps, dout, out_n);
    return st.finish();
}

int hbp_pose_pipeline(hbp_ctx* ctx, const hbp_pipeline_params* prm, const uint8_t* frames,
                      const double* M, const int* frame_idx, const float* boxes,
                      const double* height_cm, const float* joint_thr, float* kpts_img,
                      float* scores, uint32_t* ignored, float* lengths_cm, double* torso_cm,
                      void* heatmaps_out) {
    BIND(ctx);
    HBP_REQUIRE(prm && frames && M && frame_idx && boxes && height_cm && joint_thr, "null argument");
    if (!ctx->hrnet) { hbp_set_error("hbp_pose_pipeline before hbp_hrnet_load"); return HBP_ERR_STATE; }
    const int P = prm->P;
    HBP_REQUIRE(P >= 0 && prm->n_frames > 0 && prm->h > 0 && prm->w > 0, "bad shape");
    if (P == 0) return HBP_OK;
    for (int p = 0; p < P; ++p)
        HBP_REQUIRE(frame_idx[p] >= 0 && frame_idx[p] < prm->n_frames, "frame_idx out of range");
    int ih, iw, wd;
    hrnet_dims(ctx, &ih, &iw, &wd);
    const int Hh = ih / 4, Wh = iw / 4, J = 17;
    const size_t frame_bytes = (size_t)prm->n_frames * prm->h * prm->w * 3;
    uint8_t* d_frames = (uint8_t*)hbp_scratch(ctx, SC_PIPE_FRAMES, frame_bytes);
    __half* d_crops = (__half*)hbp_scratch(ctx, SC_PIPE_CROPS, (size_t)P * 3 * ih * iw * 2);
    const int hm_dtype = (heatmaps_out && prm->heatmap_dtype == HBP_F32) ? HBP_F32 : HBP_F16;
    const size_t hm_bytes = (size_t)P * J * Hh * Wh * (hm_dtype == HBP_F32 ? 4 : 2);
    void* d_hm = hbp_scratch(ctx, SC_PIPE_HM, hm_bytes);
    // one small parameter block: M (P*6 f64) | height (P f64) | boxes (P*4 f32) | thr (17 f32) | frame_idx (P i32)
    const size_t o_M = 0, o_h = o_M + (size_t)P * 6 * 8, o_b = o_h + (size_t)P * 8,
                 o_t = o_b + (size_t)P * 16, o_f = o_t + 32 * 4, par_bytes = o_f + (size_t)P * 4;
    // result block: kpts (P*17*2 f32) | scores (P*17 f32) | lengths (P*11 f32) | ignored (P u32) | torso (P f64)
    const size_t r_to = 0, r_k = r_to + (size_t)P * 8, r_s = r_k + (size_t)P * J * 2 * 4,
                 r_l = r_s + (size_t)P * J * 4, r_i = r_l + (size_t)P * 11 * 4, res_bytes = r_i + (size_t)P * 4;
    uint8_t* d_misc = (uint8_t*)hbp_scratch(ctx, SC_PIPE_MISC, par_bytes + res_bytes + 64);
    uint8_t* h_pin = (uint8_t*)hbp_pinned(ctx, par_bytes + res_bytes + 64);
    if (!d_frames || !d_crops || !d_hm || !d_misc || !h_pin) return HBP_ERR_NOMEM;
    uint8_t* d_par = d_misc;
    uint8_t* d_res = d_misc + ((par_bytes + 63) & ~size_t(63));
    uint8_t* h_par = h_pin;
    uint8_t* h_res = h_pin + ((par_bytes + 63) & ~size_t(63));
    memcpy(h_par + o_M, M, (size_t)P * 48);
    memcpy(h_par + o_h, height_cm, (size_t)P * 8);
    memcpy(h_par + o_b, boxes, (size_t)P * 16);
    memcpy(h_par + o_t, joint_thr, 17 * 4);
    memcpy(h_par + o_f, frame_idx, (size_t)P * 4);
    HBP_CUDA(cudaMemcpyAsync(d_frames, frames, frame_bytes, cudaMemcpyHostToDevice, ctx->stream));
    HBP_CUDA(cudaMemcpyAsync(d_par, h_par, par_bytes, cudaMemcpyHostToDevice, ctx->stream));
    int s = k_crop_warp(ctx, d_frames, prm->n_frames, prm->h, prm->w, (const double*)(d_par + o_M),
                        (const int*)(d_par + o_f), P, ih, iw, prm->swap_rb, d_crops, HBP_F16);
    if (s) return s;
    s = hrnet_forward(ctx, d_crops, P, d_hm, hm_dtype);
    if (s) return s;
    s = k_decode_proportions(ctx, d_hm, hm_dtype, P, J, Hh, Wh, (const float*)(d_par + o_b),
                             (const double*)(d_par + o_h), (const float*)(d_par + o_t),
                             prm->quarter_offset, nullptr, (float*)(d_res + r_k), (float*)(d_res + r_s),
                             nullptr, (uint32_t*)(d_res + r_i), (float*)(d_res + r_l), (double*)(d_res + r_to));
    if (s) return s;
    HBP_CUDA(cudaMemcpyAsync(h_res, d_res, res_bytes, cudaMemcpyDeviceToHost, ctx->stream));
    if (heatmaps_out) HBP_CUDA(cudaMemcpyAsync(heatmaps_out, d_hm, hm_bytes, cudaMemcpyDeviceToHost, ctx->stream));
    HBP_CUDA(cudaStreamSynchronize(ctx->stream));
    if (torso_cm) memcpy(torso_cm, h_res + r_to, (size_t)P * 8);
    if (kpts_img) memcpy(kpts_img, h_res + r_k, (size_t)P * J * 8);
    if (scores) memcpy(scores, h_res + r_s, (size_t)P * J * 4);
    if (lengths_cm) memcpy(lengths_cm, h_res + r_l, (size_t)P * 44);
    if (ignored) memcpy(ignored, h_res + r_i, (size_t)P * 4);
    return HBP_OK;
}

// ---- asynchronous form: two frame batches in flight --------------------------------------------
// submit(n+1) may be called before collect(n): the host->device copy of batch n+1 runs on its own
// stream while batch n is in the HRNet, so the PCIe transfer of the frames (6.2 MB for one 1080p
// frame, ~5 % of a step) leaves the critical path.  Frames, parameters and results are double
// buffered; crops and heatmaps are not (the compute stream serialises the batches).
static int pipe_reserve(uint8_t** p, size_t* cap, size_t bytes, bool pinned) {
    if (bytes <= *cap) return HBP_OK;
    if (*p) { if (pinned) cudaFreeHost(*p); else cudaFree(*p); *p = nullptr; *cap = 0; }
    const size_t n = bytes + bytes / 4 + 256;
    cudaError_t e = pinned ? cudaMallocHost((void**)p, n) : cudaMalloc((void**)p, n);
    if (e != cudaSuccess) return hbp_cuda_fail(e, "pipeline slot allocation", __FILE__, __LINE__);
    *cap = n;
    return HBP_OK;
}

int hbp_pose_pipeline_submit(hbp_ctx* ctx, const hbp_pipeline_params* prm, const uint8_t* frames,
                             const double* M, const int* frame_idx, const float* boxes,
                             const double* height_cm, const float* joint_thr, int* ticket) {
    BIND(ctx);
    HBP_REQUIRE(prm && ticket, "null argument");
    HBP_REQUIRE(prm->P == 0 || (frames && M && frame_idx && boxes && height_cm && joint_thr), "null argument");
    if (!ctx->hrnet) { hbp_set_error("hbp_pose_pipeline_submit before hbp_hrnet_load"); return HBP_ERR_STATE; }
    const int P = prm->P;
    HBP_REQUIRE(P >= 0 && prm->n_frames > 0 && prm->h > 0 && prm->w > 0, "bad shape");
    for (int p = 0; p < P; ++p)
        HBP_REQUIRE(frame_idx[p] >= 0 && frame_idx[p] < prm->n_frames, "frame_idx out of range");
    const int k = (int)(ctx->pipe_seq % HBP_PIPE_SLOTS);
    hbp_pipe_slot& sl = ctx->pipe[k];
    if (sl.busy) { hbp_set_error("pipeline slot %d has not been collected (at most %d batches in flight)", k, HBP_PIPE_SLOTS); return HBP_ERR_STATE; }
    if (P == 0) {                            // a frame without persons: nothing to enqueue, collect() returns at once
        sl.busy = true; sl.P = 0;
        *ticket = k;
        ctx->pipe_seq++;
        return HBP_OK;
    }
    if (!ctx->copy_stream) HBP_CUDA(cudaStreamCreateWithFlags(&ctx->copy_stream, cudaStreamNonBlocking));
    if (!sl.ev_h2d) {
        HBP_CUDA(cudaEventCreateWithFlags(&sl.ev_h2d, cudaEventDisableTiming));
        HBP_CUDA(cudaEventCreateWithFlags(&sl.ev_crop, cudaEventDisableTiming));
        HBP_CUDA(cudaEventCreateWithFlags(&sl.ev_done, cudaEventDisableTiming));
    }
    int ih, iw, wd;
    hrnet_dims(ctx, &ih, &iw, &wd);
    const int Hh = ih / 4, Wh = iw / 4, J = 17;
    const size_t frame_bytes = (size_t)prm->n_frames * prm->h * prm->w * 3;
    const size_t o_M = 0, o_h = o_M + (size_t)P * 6 * 8, o_b = o_h + (size_t)P * 8,
                 o_t = o_b + (size_t)P * 16, o_f = o_t + 32 * 4, par_bytes = o_f + (size_t)P * 4;
    const size_t r_to = 0, r_k = r_to + (size_t)P * 8, r_s = r_k + (size_t)P * J * 2 * 4,
                 r_l = r_s + (size_t)P * J * 4, r_i = r_l + (size_t)P * 11 * 4, res_bytes = r_i + (size_t)P * 4;
    const size_t par_pad = (par_bytes + 63) & ~size_t(63);
    // (re)allocation only while nothing of this slot is in flight: its previous batch was collected
    if (frame_bytes > sl.frames_cap || par_pad + res_bytes > sl.misc_cap) HBP_CUDA(cudaDeviceSynchronize());
    int st = pipe_reserve(&sl.d_frames, &sl.frames_cap, frame_bytes, false);
    if (!st) st = pipe_reserve(&sl.d_misc, &sl.misc_cap, par_pad + res_bytes, false);
    if (!st) st = pipe_reserve(&sl.h_pin, &sl.pin_cap, par_pad + res_bytes, true);
    if (st) return st;
    __half* d_crops = (__half*)hbp_scratch(ctx, SC_PIPE_CROPS, (size_t)P * 3 * ih * iw * 2);
    void* d_hm = hbp_scratch(ctx, SC_PIPE_HM, (size_t)P * J * Hh * Wh * 2);
    if (!d_crops || !d_hm) return HBP_ERR_NOMEM;
    uint8_t* d_par = sl.d_misc;
    uint8_t* d_res = sl.d_misc + par_pad;
    uint8_t* h_par = sl.h_pin;
    uint8_t* h_res = sl.h_pin + par_pad;
    memcpy(h_par + o_M, M, (size_t)P * 48);
    memcpy(h_par + o_h, height_cm, (size_t)P * 8);
    memcpy(h_par + o_b, boxes, (size_t)P * 16);
    memcpy(h_par + o_t, joint_thr, 17 * 4);
    memcpy(h_par + o_f, frame_idx, (size_t)P * 4);
    // copy stream: the slot's frame buffer is free once the crop kernel of its previous batch has read it
    if (sl.crop_recorded) HBP_CUDA(cudaStreamWaitEvent(ctx->copy_stream, sl.ev_crop, 0));
    HBP_CUDA(cudaMemcpyAsync(sl.d_frames, frames, frame_bytes, cudaMemcpyHostToDevice, ctx->copy_stream));
    HBP_CUDA(cudaMemcpyAsync(d_par, h_par, par_bytes, cudaMemcpyHostToDevice, ctx->copy_stream));
    HBP_CUDA(cudaEventRecord(sl.ev_h2d, ctx->copy_stream));
    // compute stream
    HBP_CUDA(cudaStreamWaitEvent(ctx->stream, sl.ev_h2d, 0));
    int s = k_crop_warp(ctx, sl.d_frames, prm->n_frames, prm->h, prm->w, (const double*)(d_par + o_M),
                        (const int*)(d_par + o_f), P, ih, iw, prm->swap_rb, d_crops, HBP_F16);
    if (s) return s;
    HBP_CUDA(cudaEventRecord(sl.ev_crop, ctx->stream));
    sl.crop_recorded = true;
    s = hrnet_forward(ctx, d_crops, P, d_hm, HBP_F16);
    if (s) return s;
    s = k_decode_proportions(ctx, d_hm, HBP_F16, P, J, Hh, Wh, (const float*)(d_par + o_b),
                             (const double*)(d_par + o_h), (const float*)(d_par + o_t),
                             prm->quarter_offset, nullptr, (float*)(d_res + r_k), (float*)(d_res + r_s),
                             nullptr, (uint32_t*)(d_res + r_i), (float*)(d_res + r_l), (double*)(d_res + r_to));
    if (s) return s;
    HBP_CUDA(cudaMemcpyAsync(h_res, d_res, res_bytes, cudaMemcpyDeviceToHost, ctx->stream));
    HBP_CUDA(cudaEventRecord(sl.ev_done, ctx->stream));
    sl.busy = true; sl.P = P; sl.par_bytes = par_pad; sl.res_bytes = res_bytes;
    *ticket = k;
    ctx->pipe_seq++;
    return HBP_OK;
}

int hbp_pose_pipeline_collect(hbp_ctx* ctx, int ticket, float* kpts_img, float* scores, uint32_t* ignored,
                              float* lengths_cm, double* torso_cm) {
    BIND(ctx);
    HBP_REQUIRE(ticket >= 0 && ticket < HBP_PIPE_SLOTS, "bad ticket");
    hbp_pipe_slot& sl = ctx->pipe[ticket];
    if (!sl.busy || sl.det_pose) { hbp_set_error("ticket %d has no pose batch in flight", ticket); return HBP_ERR_STATE; }
    if (sl.P == 0) { sl.busy = false; return HBP_OK; }
    HBP_CUDA(cudaEventSynchronize(sl.ev_done));
    const int P = sl.P, J = 17;
    const size_t r_to = 0, r_k = r_to + (size_t)P * 8, r_s = r_k + (size_t)P * J * 2 * 4,
                 r_l = r_s + (size_t)P * J * 4, r_i = r_l + (size_t)P * 11 * 4;
    const uint8_t* h_res = sl.h_pin + sl.par_bytes;
    if (torso_cm) memcpy(torso_cm, h_res + r_to, (size_t)P * 8);
    if (kpts_img) memcpy(kpts_img, h_res + r_k, (size_t)P * J * 8);
    if (scores) memcpy(scores, h_res + r_s, (size_t)P * J * 4);
    if (lengths_cm) memcpy(lengths_cm, h_res + r_l, (size_t)P * 44);
    if (ignored) memcpy(ignored, h_res + r_i, (size_t)P * 4);
    sl.busy = false;
    return HBP_OK;
}

}  // extern "C"
