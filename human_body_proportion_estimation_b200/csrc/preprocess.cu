// K1: frame preprocessing from NHWC uint8.
//
// Reference call sites (relative to the reference's human_body_length_est/):
//   COPY ....... person_det_pose_edet4_trtserver.py:15-18 (BGR->RGB, no resize with
//                the shipped dynamic-shape ensemble, uint8 out)
//   STRETCH .... modules/pose_estimator.py:29-45, pose_est_hrnet_trtserver.py:15-19
//                (cv2.resize to the model size, /255, CHW float)
//   LETTERBOX .. obj_det_yolov5_onnx.py:27-36 + modules/onnx_utils.py:225-235
//                (aspect-keeping resize pasted centred on grey 128, /255, CHW)
//
// The resampler is cv2.resize(INTER_LINEAR) for uint8, reproduced bit for bit
// (OpenCV resize.cpp; restated in oracle/imgproc.py:resize_linear_u8_cv2):
//     fx = float((dx+0.5)*scale - 0.5); sx = floor(fx); fx -= sx
//     horizontally a clamped left tap zeroes fx; vertically only the row indices clamp
//     a0 = rint((1-fx)*2048), a1 = rint(fx*2048)   (int16 coefficients)
//     r  = S[sx]*a0 + S[sx+1]*a1                     (per source row)
//     v  = (((b0*(r0>>4))>>16) + ((b1*(r1>>4))>>16) + 2) >> 2
// LETTERBOX_PIL reproduces the reference's own resampler, PIL's antialiased BICUBIC
// (modules/onnx_utils.py:232; Pillow src/libImaging/Resample.c, restated in
// oracle/imgproc.py:resize_bicubic_pil): per output column / row a window of
// ceil(2*scale)*2+1 taps with 22-bit fixed-point coefficients (computed here on the host in the
// same double arithmetic, uploaded as tables), a horizontal pass into a uint8 intermediate,
// then the vertical pass -- both round with 2^21 and clip to [0,255].
// The COPY/uint8/NHWC case is a pure stream: each thread moves 16 pixels with
// three 16-byte loads and three 16-byte stores (byte permute in registers).
// HBM-bound: H*W*3 bytes read + 3*out_h*out_w*e bytes written per frame.
#include "hbp_internal.cuh"
#include <algorithm>
#include <cmath>
#include <vector>

namespace {

// ---- COPY, u8 NHWC -> u8 NHWC, optional channel reversal -------------------
__global__ void __launch_bounds__(256)
swap_copy_kernel(const uint8_t* __restrict__ in, uint8_t* __restrict__ out, size_t n_pix, int swap_rb) {
    const size_t n_grp = n_pix / 16;                 // 16 pixels = 48 bytes = 3 x uint4
    const size_t stride = (size_t)gridDim.x * blockDim.x;
    for (size_t g = (size_t)blockIdx.x * blockDim.x + threadIdx.x; g < n_grp; g += stride) {
        const uint4* src = reinterpret_cast<const uint4*>(in + g * 48);
        uint4 v[3] = {__ldg(src), __ldg(src + 1), __ldg(src + 2)};
        if (swap_rb) {
            uint8_t* b = reinterpret_cast<uint8_t*>(v);
#pragma unroll
            for (int k = 0; k < 16; ++k) {
                const uint8_t t = b[3 * k];
                b[3 * k] = b[3 * k + 2];
                b[3 * k + 2] = t;
            }
        }
        uint4* dst = reinterpret_cast<uint4*>(out + g * 48);
        dst[0] = v[0]; dst[1] = v[1]; dst[2] = v[2];
    }
    // tail pixels
    const size_t tail0 = n_grp * 16;
    for (size_t p = tail0 + (size_t)blockIdx.x * blockDim.x + threadIdx.x; p < n_pix; p += stride) {
        const uint8_t a = in[3 * p], b = in[3 * p + 1], c = in[3 * p + 2];
        out[3 * p] = swap_rb ? c : a; out[3 * p + 1] = b; out[3 * p + 2] = swap_rb ? a : c;
    }
}

struct AxisTap { int i0, i1, c0, c1; };

__device__ __forceinline__ AxisTap axis_tap(int d, double scale, int src_n, bool vertical) {
    float f = (float)__dsub_rn(__dmul_rn((double)d + 0.5, scale), 0.5);
    int s = (int)floorf(f);
    f = __fsub_rn(f, (float)s);
    AxisTap t;
    if (vertical) {
        t.i0 = min(max(s, 0), src_n - 1);
        t.i1 = min(max(s + 1, 0), src_n - 1);
    } else {
        if (s < 0) { f = 0.f; s = 0; }
        if (s >= src_n - 1) { f = 0.f; s = src_n - 1; }
        t.i0 = s;
        t.i1 = min(s + 1, src_n - 1);
    }
    t.c0 = __float2int_rn(__fmul_rn(__fsub_rn(1.f, f), 2048.f));
    t.c1 = __float2int_rn(__fmul_rn(f, 2048.f));
    return t;
}

template <typename T> __device__ __forceinline__ T cvt(int v);
template <> __device__ __forceinline__ uint8_t cvt<uint8_t>(int v) { return (uint8_t)v; }
template <> __device__ __forceinline__ float cvt<float>(int v) { return __fdiv_rn((float)v, 255.f); }
template <> __device__ __forceinline__ __half cvt<__half>(int v) { return __float2half_rn(__fdiv_rn((float)v, 255.f)); }

// ---- generic: resize (STRETCH / LETTERBOX) or identity (COPY) ---------------
// content rectangle [ox,ox+nw) x [oy,oy+nh) of the output is the resized frame,
// the rest is pad_value.  One thread = 4 consecutive output pixels of one row.
template <typename T>
__global__ void __launch_bounds__(256)
resize_kernel(const uint8_t* __restrict__ in, int n, int H, int W, T* __restrict__ out, int out_h,
              int out_w, int ox, int oy, int nw, int nh, double scale_x, double scale_y,
              int identity, int swap_rb, int pad_value, int nchw) {
    const int groups = (out_w + 3) >> 2;
    const size_t total = (size_t)n * out_h * groups;
    const size_t stride = (size_t)gridDim.x * blockDim.x;
    for (size_t t = (size_t)blockIdx.x * blockDim.x + threadIdx.x; t < total; t += stride) {
        const int gx = (int)(t % groups);
        const int y = (int)((t / groups) % out_h);
        const int f = (int)(t / ((size_t)groups * out_h));
        const uint8_t* __restrict__ src = in + (size_t)f * H * W * 3;
        const bool row_in = y >= oy && y < oy + nh;
        AxisTap ty{0, 0, 0, 0};
        if (row_in && !identity) ty = axis_tap(y - oy, scale_y, H, true);
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            const int x = gx * 4 + k;
            if (x >= out_w) break;
            int v[3] = {pad_value, pad_value, pad_value};
            if (row_in && x >= ox && x < ox + nw) {
                if (identity) {
                    const uint8_t* q = src + ((size_t)(y - oy) * W + (x - ox)) * 3;
                    v[0] = __ldg(q); v[1] = __ldg(q + 1); v[2] = __ldg(q + 2);
                } else {
                    const AxisTap tx = axis_tap(x - ox, scale_x, W, false);
                    const uint8_t* r0 = src + (size_t)ty.i0 * W * 3;
                    const uint8_t* r1 = src + (size_t)ty.i1 * W * 3;
#pragma unroll
                    for (int c = 0; c < 3; ++c) {
                        const int h0 = (int)__ldg(r0 + tx.i0 * 3 + c) * tx.c0 + (int)__ldg(r0 + tx.i1 * 3 + c) * tx.c1;
                        const int h1 = (int)__ldg(r1 + tx.i0 * 3 + c) * tx.c0 + (int)__ldg(r1 + tx.i1 * 3 + c) * tx.c1;
                        int r = (((ty.c0 * (h0 >> 4)) >> 16) + ((ty.c1 * (h1 >> 4)) >> 16) + 2) >> 2;
                        v[c] = min(max(r, 0), 255);
                    }
                }
            }
            const int c0 = swap_rb ? v[2] : v[0], c2 = swap_rb ? v[0] : v[2];
            if (nchw) {
                const size_t plane = (size_t)out_h * out_w;
                T* o = out + (size_t)f * 3 * plane + (size_t)y * out_w + x;
                o[0] = cvt<T>(c0); o[plane] = cvt<T>(v[1]); o[2 * plane] = cvt<T>(c2);
            } else {
                T* o = out + (((size_t)f * out_h + y) * out_w + x) * 3;
                o[0] = cvt<T>(c0); o[1] = cvt<T>(v[1]); o[2] = cvt<T>(c2);
            }
        }
    }
}

template <typename T>
int launch_resize(hbp_ctx* ctx, const uint8_t* in, int n, int H, int W, void* out, int out_h, int out_w,
                  int ox, int oy, int nw, int nh, double sx, double sy, int identity, int swap_rb,
                  int pad_value, int nchw) {
    const size_t total = (size_t)n * out_h * ((out_w + 3) / 4);
    const int blocks = (int)std::min<size_t>((total + 255) / 256, (size_t)ctx->sm_count * 16);
    resize_kernel<T><<<blocks, 256, 0, ctx->stream>>>(in, n, H, W, (T*)out, out_h, out_w, ox, oy, nw, nh,
                                                      sx, sy, identity, swap_rb, pad_value, nchw);
    HBP_LAUNCH_CHECK(ctx);
    return HBP_OK;
}

// ---- PIL antialiased bicubic (two 8-bit passes) ------------------------------
constexpr int kPilBits = 32 - 8 - 2;

__device__ __forceinline__ int pil_clip8(int v) { return min(max(v >> kPilBits, 0), 255); }

// horizontal pass: rows [row0, row0+rows) of every frame -> tmp (n, rows, nw, 3) u8
__global__ void __launch_bounds__(256)
pil_hpass_kernel(const uint8_t* __restrict__ in, int n, int H, int W, int row0, int rows, int nw,
                 const int* __restrict__ bounds, const int* __restrict__ kk, int ksize, uint8_t* __restrict__ tmp) {
    const size_t total = (size_t)n * rows * nw;
    const size_t stride = (size_t)gridDim.x * blockDim.x;
    for (size_t t = (size_t)blockIdx.x * blockDim.x + threadIdx.x; t < total; t += stride) {
        const int xx = (int)(t % nw);
        const int r = (int)((t / nw) % rows);
        const int f = (int)(t / ((size_t)nw * rows));
        const int xmin = __ldg(bounds + 2 * xx), xn = __ldg(bounds + 2 * xx + 1);
        const int* k = kk + (size_t)xx * ksize;
        const uint8_t* src = in + (((size_t)f * H + row0 + r) * W + xmin) * 3;
        int s0 = 1 << (kPilBits - 1), s1 = s0, s2 = s0;
        for (int x = 0; x < xn; ++x) {
            const int c = __ldg(k + x);
            s0 += (int)__ldg(src + 3 * x) * c;
            s1 += (int)__ldg(src + 3 * x + 1) * c;
            s2 += (int)__ldg(src + 3 * x + 2) * c;
        }
        uint8_t* o = tmp + t * 3;
        o[0] = (uint8_t)pil_clip8(s0); o[1] = (uint8_t)pil_clip8(s1); o[2] = (uint8_t)pil_clip8(s2);
    }
}

// vertical pass + paste on the pad canvas + /255 + layout
template <typename T>
__global__ void __launch_bounds__(256)
pil_vpass_kernel(const uint8_t* __restrict__ tmp, int n, int rows, int row0, int nw, int nh,
                 const int* __restrict__ bounds, const int* __restrict__ kk, int ksize, T* __restrict__ out,
                 int out_h, int out_w, int ox, int oy, int swap_rb, int pad_value, int nchw, int vertical) {
    const size_t total = (size_t)n * out_h * out_w;
    const size_t stride = (size_t)gridDim.x * blockDim.x;
    for (size_t t = (size_t)blockIdx.x * blockDim.x + threadIdx.x; t < total; t += stride) {
        const int x = (int)(t % out_w);
        const int y = (int)((t / out_w) % out_h);
        const int f = (int)(t / ((size_t)out_w * out_h));
        int v[3] = {pad_value, pad_value, pad_value};
        if (y >= oy && y < oy + nh && x >= ox && x < ox + nw) {
            const int yy = y - oy, xx = x - ox;
            if (vertical) {
                const int ymin = __ldg(bounds + 2 * yy) - row0, yn = __ldg(bounds + 2 * yy + 1);
                const int* k = kk + (size_t)yy * ksize;
                const uint8_t* src = tmp + (((size_t)f * rows + ymin) * nw + xx) * 3;
                int s0 = 1 << (kPilBits - 1), s1 = s0, s2 = s0;
                for (int r = 0; r < yn; ++r) {
                    const int c = __ldg(k + r);
                    const uint8_t* q = src + (size_t)r * nw * 3;
                    s0 += (int)__ldg(q) * c; s1 += (int)__ldg(q + 1) * c; s2 += (int)__ldg(q + 2) * c;
                }
                v[0] = pil_clip8(s0); v[1] = pil_clip8(s1); v[2] = pil_clip8(s2);
            } else {
                const uint8_t* q = tmp + (((size_t)f * rows + yy - row0) * nw + xx) * 3;
                v[0] = __ldg(q); v[1] = __ldg(q + 1); v[2] = __ldg(q + 2);
            }
        }
        const int c0 = swap_rb ? v[2] : v[0], c2 = swap_rb ? v[0] : v[2];
        if (nchw) {
            const size_t plane = (size_t)out_h * out_w;
            T* o = out + (size_t)f * 3 * plane + (size_t)y * out_w + x;
            o[0] = cvt<T>(c0); o[plane] = cvt<T>(v[1]); o[2 * plane] = cvt<T>(c2);
        } else {
            T* o = out + t * 3;
            o[0] = cvt<T>(c0); o[1] = cvt<T>(v[1]); o[2] = cvt<T>(c2);
        }
    }
}

// Pillow Resample.c precompute_coeffs + normalize_coeffs_8bpc for the BICUBIC filter (support 2, a = -0.5),
// box = the whole axis.  Same double operations in the same order as the C source.
static double pil_bicubic(double x) {
    const double a = -0.5;
    if (x < 0.0) x = -x;
    if (x < 1.0) return ((a + 2.0) * x - (a + 3.0)) * x * x + 1;
    if (x < 2.0) return (((x - 5) * x + 8) * x - 4) * a;
    return 0.0;
}

static int pil_coeffs(int in_size, int out_size, std::vector<int>& bounds, std::vector<int>& kk) {
    double scale, filterscale;
    filterscale = scale = (double)((float)in_size - 0.0f) / out_size;
    if (filterscale < 1.0) filterscale = 1.0;
    const double support = 2.0 * filterscale;
    const int ksize = (int)ceil(support) * 2 + 1;
    bounds.assign((size_t)out_size * 2, 0);
    kk.assign((size_t)out_size * ksize, 0);
    std::vector<double> w(ksize);
    const double ss = 1.0 / filterscale;
    for (int xx = 0; xx < out_size; ++xx) {
        const double center = 0.0f + (xx + 0.5) * scale;
        double ww = 0.0;
        int xmin = (int)(center - support + 0.5);
        if (xmin < 0) xmin = 0;
        int xmax = (int)(center + support + 0.5);
        if (xmax > in_size) xmax = in_size;
        xmax -= xmin;
        for (int x = 0; x < xmax; ++x) {
            w[x] = pil_bicubic((x + xmin - center + 0.5) * ss);
            ww += w[x];
        }
        for (int x = 0; x < xmax; ++x) {
            if (ww != 0.0) w[x] /= ww;
            kk[(size_t)xx * ksize + x] = w[x] < 0 ? (int)(-0.5 + w[x] * (1 << kPilBits)) : (int)(0.5 + w[x] * (1 << kPilBits));
        }
        bounds[2 * xx] = xmin;
        bounds[2 * xx + 1] = xmax;
    }
    return ksize;
}

template <typename T>
int launch_pil(hbp_ctx* ctx, const uint8_t* in, int n, int H, int W, void* out, int out_h, int out_w, int ox, int oy,
               int nw, int nh, int swap_rb, int pad_value, int nchw) {
    // Resample.c ImagingResampleInner: horizontal pass over the source rows the vertical pass needs, then vertical
    const bool need_h = nw != W, need_v = nh != H;
    std::vector<int> bh, kh, bv, kv;
    const int ks_h = pil_coeffs(W, nw, bh, kh);
    const int ks_v = pil_coeffs(H, nh, bv, kv);
    int row0 = 0, rows = H;
    if (need_v) { row0 = bv[0]; rows = bv[2 * (nh - 1)] + bv[2 * (nh - 1) + 1] - row0; }
    const size_t n_tab = bh.size() + kh.size() + bv.size() + kv.size();
    int* d_tab = (int*)hbp_scratch(ctx, SC_PRE_COEF, n_tab * sizeof(int));
    if (!d_tab) return HBP_ERR_NOMEM;
    std::vector<int> tab;
    tab.reserve(n_tab);
    tab.insert(tab.end(), bh.begin(), bh.end()); tab.insert(tab.end(), kh.begin(), kh.end());
    tab.insert(tab.end(), bv.begin(), bv.end()); tab.insert(tab.end(), kv.begin(), kv.end());
    HBP_CUDA(cudaMemcpyAsync(d_tab, tab.data(), n_tab * sizeof(int), cudaMemcpyHostToDevice, ctx->stream));
    HBP_CUDA(cudaStreamSynchronize(ctx->stream));             // `tab` is a local
    const int* d_bh = d_tab; const int* d_kh = d_bh + bh.size();
    const int* d_bv = d_kh + kh.size(); const int* d_kv = d_bv + bv.size();
    const uint8_t* tmp = in;
    int t_rows = H, t_row0 = 0;
    if (need_h) {
        uint8_t* d_tmp = (uint8_t*)hbp_scratch(ctx, SC_PRE_TMP, (size_t)n * rows * nw * 3);
        if (!d_tmp) return HBP_ERR_NOMEM;
        const size_t total = (size_t)n * rows * nw;
        const int blocks = (int)std::min<size_t>((total + 255) / 256, (size_t)ctx->sm_count * 16);
        pil_hpass_kernel<<<blocks, 256, 0, ctx->stream>>>(in, n, H, W, row0, rows, nw, d_bh, d_kh, ks_h, d_tmp);
        HBP_LAUNCH_CHECK(ctx);
        tmp = d_tmp; t_rows = rows; t_row0 = row0;
    }
    const size_t total = (size_t)n * out_h * out_w;
    const int blocks = (int)std::min<size_t>((total + 255) / 256, (size_t)ctx->sm_count * 16);
    pil_vpass_kernel<T><<<blocks, 256, 0, ctx->stream>>>(tmp, n, t_rows, t_row0, nw, nh, d_bv, d_kv, ks_v, (T*)out, out_h, out_w,
                                                         ox, oy, swap_rb, pad_value, nchw, need_v ? 1 : 0);
    HBP_LAUNCH_CHECK(ctx);
    return HBP_OK;
}

}  // namespace

int k_preprocess(hbp_ctx* ctx, const uint8_t* frames, int n, int h, int w, int mode, int out_h,
                 int out_w, int swap_rb, int pad_value, void* out, int out_dtype, int out_layout) {
    if (mode == HBP_PRE_COPY && out_dtype == HBP_U8 && out_layout == HBP_NHWC &&
        ((reinterpret_cast<uintptr_t>(frames) | reinterpret_cast<uintptr_t>(out)) & 15) == 0) {
        const size_t n_pix = (size_t)n * h * w;
        const int blocks = (int)std::min<size_t>((n_pix / 16 + 255) / 256 + 1, (size_t)ctx->sm_count * 8);
        swap_copy_kernel<<<blocks, 256, 0, ctx->stream>>>(frames, (uint8_t*)out, n_pix, swap_rb);
        HBP_LAUNCH_CHECK(ctx);
        return HBP_OK;
    }
    int ox = 0, oy = 0, nw = out_w, nh = out_h, identity = 0;
    if (mode == HBP_PRE_COPY) {
        identity = 1;
    } else if (mode == HBP_PRE_LETTERBOX || mode == HBP_PRE_LETTERBOX_PIL) {
        // onnx_utils.py:225-235: scale = min(w/iw, h/ih); nw = int(iw*scale); paste at ((w-nw)//2, (h-nh)//2)
        const double a = (double)out_w / (double)w, b = (double)out_h / (double)h;
        const double scale = a < b ? a : b;
        nw = (int)((double)w * scale);
        nh = (int)((double)h * scale);
        if (nw < 1 || nh < 1) { hbp_set_error("letterbox target too small"); return HBP_ERR_INVALID; }
        ox = (out_w - nw) / 2;
        oy = (out_h - nh) / 2;
    }
    const int nchw = out_layout == HBP_NCHW;
    if (mode == HBP_PRE_LETTERBOX_PIL) {
        if (out_dtype == HBP_U8) return launch_pil<uint8_t>(ctx, frames, n, h, w, out, out_h, out_w, ox, oy, nw, nh, swap_rb, pad_value, nchw);
        if (out_dtype == HBP_F16) return launch_pil<__half>(ctx, frames, n, h, w, out, out_h, out_w, ox, oy, nw, nh, swap_rb, pad_value, nchw);
        return launch_pil<float>(ctx, frames, n, h, w, out, out_h, out_w, ox, oy, nw, nh, swap_rb, pad_value, nchw);
    }
    // cv2: scale_x = 1. / (dsize.width / (double)ssize.width)
    const double sx = 1.0 / ((double)nw / (double)w), sy = 1.0 / ((double)nh / (double)h);
    if (out_dtype == HBP_U8)
        return launch_resize<uint8_t>(ctx, frames, n, h, w, out, out_h, out_w, ox, oy, nw, nh, sx, sy, identity, swap_rb, pad_value, nchw);
    if (out_dtype == HBP_F16)
        return launch_resize<__half>(ctx, frames, n, h, w, out, out_h, out_w, ox, oy, nw, nh, sx, sy, identity, swap_rb, pad_value, nchw);
    return launch_resize<float>(ctx, frames, n, h, w, out, out_h, out_w, ox, oy, nw, nh, sx, sy, identity, swap_rb, pad_value, nchw);
}
