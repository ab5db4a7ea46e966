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
    // per CTA, once: the horizontal taps of every content column (byte offsets of the two source pixels and the two
    // 11-bit coefficients: the double-precision coordinate arithmetic used to run per output pixel) and the 256
    // possible output values (v / 255 rounded like the reference's float division: three IEEE divisions per pixel
    // otherwise).  The kernel was instruction-issue bound: 51 % of HBM for 4K -> 640.
    extern __shared__ __align__(16) unsigned char s_raw[];
    int4* s_tx = reinterpret_cast<int4*>(s_raw);                         // nw entries: {3*i0, 3*i1, c0, c1}
    T* s_cvt = reinterpret_cast<T*>(s_raw + (size_t)((nw + 3) & ~3) * sizeof(int4));   // 256 entries
    for (int v = threadIdx.x; v < 256; v += blockDim.x) s_cvt[v] = cvt<T>(v);
    if (!identity)
        for (int x = threadIdx.x; x < nw; x += blockDim.x) {
            const AxisTap tx = axis_tap(x, scale_x, W, false);
            s_tx[x] = make_int4(tx.i0 * 3, tx.i1 * 3, tx.c0, tx.c1);
        }
    __syncthreads();
    const int groups = (out_w + 3) >> 2;
    const size_t total = (size_t)n * out_h * groups;
    const size_t stride = (size_t)gridDim.x * blockDim.x;
    const int row_bytes = W * 3;
    const bool packed = nchw && (out_w & 3) == 0 && (reinterpret_cast<uintptr_t>(out) & (4 * sizeof(T) - 1)) == 0;   // four pixels of a plane row = one aligned store
    for (size_t t = (size_t)blockIdx.x * blockDim.x + threadIdx.x; t < total; t += stride) {
        const int gx = (int)(t % groups);
        const int y = (int)((t / groups) % out_h);
        const int f = (int)(t / ((size_t)groups * out_h));
        const uint8_t* __restrict__ src = in + (size_t)f * H * W * 3;
        const bool row_in = y >= oy && y < oy + nh;
        AxisTap ty{0, 0, 0, 0};
        if (row_in && !identity) ty = axis_tap(y - oy, scale_y, H, true);
        const uint8_t* r0 = src + (size_t)ty.i0 * W * 3;
        const uint8_t* r1 = src + (size_t)ty.i1 * W * 3;
        int vv[4][3];
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            const int x = gx * 4 + k;
            vv[k][0] = vv[k][1] = vv[k][2] = pad_value;
            if (x < out_w && row_in && x >= ox && x < ox + nw) {
                if (identity) {
                    const uint8_t* q = src + ((size_t)(y - oy) * W + (x - ox)) * 3;
                    vv[k][0] = __ldg(q); vv[k][1] = __ldg(q + 1); vv[k][2] = __ldg(q + 2);
                } else {
                    const int4 tx = s_tx[x - ox];
                    if (tx.x + 12 <= row_bytes) {
                        // the two taps are six consecutive bytes: three aligned 32-bit loads per source row and two funnel
                        // shifts instead of twelve byte loads (the kernel was bound by L1 requests, not by DRAM bytes);
                        // the loads stay inside the row, the last two columns of a row take the byte path below
                        int a0[3], b0[3], a1[3], b1[3];
                        auto fetch = [&](const uint8_t* row, int (&pa)[3], int (&pb)[3]) {
                            const uintptr_t ad = reinterpret_cast<uintptr_t>(row + tx.x);
                            const uint32_t* w = reinterpret_cast<const uint32_t*>(ad & ~(uintptr_t)3);
                            const uint32_t sh = (uint32_t)(ad & 3) * 8u;
                            const uint32_t w0 = __ldg(w), w1 = __ldg(w + 1), w2 = __ldg(w + 2);
                            const uint32_t lo = __funnelshift_r(w0, w1, sh), hi = __funnelshift_r(w1, w2, sh);
                            pa[0] = lo & 0xff; pa[1] = (lo >> 8) & 0xff; pa[2] = (lo >> 16) & 0xff;
                            pb[0] = lo >> 24; pb[1] = hi & 0xff; pb[2] = (hi >> 8) & 0xff;
                        };
                        fetch(r0, a0, b0);
                        fetch(r1, a1, b1);
#pragma unroll
                        for (int c = 0; c < 3; ++c) {
                            const int h0 = a0[c] * tx.z + b0[c] * tx.w;
                            const int h1 = a1[c] * tx.z + b1[c] * tx.w;
                            const int r = (((ty.c0 * (h0 >> 4)) >> 16) + ((ty.c1 * (h1 >> 4)) >> 16) + 2) >> 2;
                            vv[k][c] = min(max(r, 0), 255);
                        }
                    } else {
#pragma unroll
                        for (int c = 0; c < 3; ++c) {
                            const int h0 = (int)__ldg(r0 + tx.x + c) * tx.z + (int)__ldg(r0 + tx.y + c) * tx.w;
                            const int h1 = (int)__ldg(r1 + tx.x + c) * tx.z + (int)__ldg(r1 + tx.y + c) * tx.w;
                            const int r = (((ty.c0 * (h0 >> 4)) >> 16) + ((ty.c1 * (h1 >> 4)) >> 16) + 2) >> 2;
                            vv[k][c] = min(max(r, 0), 255);
                        }
                    }
                }
            }
            if (swap_rb) { const int tmp = vv[k][0]; vv[k][0] = vv[k][2]; vv[k][2] = tmp; }
        }
        if (packed) {
            const size_t plane = (size_t)out_h * out_w;
            T* o = out + (size_t)f * 3 * plane + (size_t)y * out_w + gx * 4;
#pragma unroll
            for (int c = 0; c < 3; ++c) {
                struct alignas(4 * sizeof(T)) Pack { T v[4]; } pk;
#pragma unroll
                for (int k = 0; k < 4; ++k) pk.v[k] = s_cvt[vv[k][c]];
                *reinterpret_cast<Pack*>(o + c * plane) = pk;
            }
        } else {
#pragma unroll
            for (int k = 0; k < 4; ++k) {
                const int x = gx * 4 + k;
                if (x >= out_w) break;
                if (nchw) {
                    const size_t plane = (size_t)out_h * out_w;
                    T* o = out + (size_t)f * 3 * plane + (size_t)y * out_w + x;
                    o[0] = s_cvt[vv[k][0]]; o[plane] = s_cvt[vv[k][1]]; o[2 * plane] = s_cvt[vv[k][2]];
                } else {
                    T* o = out + (((size_t)f * out_h + y) * out_w + x) * 3;
                    o[0] = s_cvt[vv[k][0]]; o[1] = s_cvt[vv[k][1]]; o[2] = s_cvt[vv[k][2]];
                }
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
    const size_t smem = (size_t)((nw + 3) & ~3) * sizeof(int4) + 256 * sizeof(T);
    if (smem > 48 * 1024) {
        const uint32_t flag = sizeof(T) == 1 ? ATTR_RESIZE_U8 : sizeof(T) == 2 ? ATTR_RESIZE_F16 : ATTR_RESIZE_F32;   // per device
        if (!(ctx->attr_flags & flag)) {
            HBP_CUDA(cudaFuncSetAttribute(resize_kernel<T>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
            ctx->attr_flags |= flag;
        }
        if (smem > 200 * 1024) { hbp_set_error("resize: output rows wider than %d pixels are not supported", (200 * 1024 - 1024) / 16); return HBP_ERR_INVALID; }
    }
    resize_kernel<T><<<blocks, 256, smem, ctx->stream>>>(in, n, H, W, (T*)out, out_h, out_w, ox, oy, nw, nh,
                                                         sx, sy, identity, swap_rb, pad_value, nchw);
    HBP_LAUNCH_CHECK(ctx);
    return HBP_OK;
}

// ---- PIL antialiased bicubic (two 8-bit passes) ------------------------------
constexpr int kPilBits = 32 - 8 - 2;

__device__ __forceinline__ int pil_clip8(int v) { return min(max(v >> kPilBits, 0), 255); }

// horizontal pass: rows [row0, row0+rows) of every frame -> tmp (n, rows, nw, 3) u8.
// One CTA = kPilRows consecutive source rows of one frame: the rows are staged in shared memory with 16-byte
// loads (a tap window is 3*(2*scale*2+1) scattered bytes per output -- 75 at 4K -> 640 -- which as global byte
// loads made the kernel 7 % of the HBM roofline), the coefficient table is stored tap-major so that a warp
// reads one coalesced line per tap, and each coefficient is applied to all staged rows.
constexpr int kPilRows = 2;

__global__ void __launch_bounds__(256)
pil_hpass_kernel(const uint8_t* __restrict__ in, int n, int H, int W, int row0, int rows, int nw,
                 const int* __restrict__ bounds, const int* __restrict__ kk_t, int ksize, uint8_t* __restrict__ tmp) {
    extern __shared__ __align__(16) uint8_t s_rows[];          // kPilRows x pitch, each row keeps its global 16-byte phase
    const int groups = (rows + kPilRows - 1) / kPilRows;
    const int f = blockIdx.x / groups, r0 = (blockIdx.x % groups) * kPilRows;
    const int nr = min(kPilRows, rows - r0);
    const int row_bytes = W * 3;
    const int pitch = ((row_bytes + 15 + 15) / 16) * 16;
    const uint8_t* frames_end = in + (size_t)n * H * W * 3;
    int ph[kPilRows];
#pragma unroll
    for (int r = 0; r < kPilRows; ++r) {
        ph[r] = 0;
        if (r >= nr) continue;
        const uint8_t* g0 = in + ((size_t)f * H + row0 + r0 + r) * row_bytes;
        ph[r] = (int)(reinterpret_cast<uintptr_t>(g0) & 15);
        const uint8_t* ga0 = g0 - ph[r];
        for (int c = threadIdx.x; c < pitch / 16; c += blockDim.x) {
            const uint8_t* ga = ga0 + 16 * c;
            uint4 v = make_uint4(0, 0, 0, 0);
            if (ga >= in && ga + 16 <= frames_end) v = __ldg(reinterpret_cast<const uint4*>(ga));
            else {
                uint8_t* vb = reinterpret_cast<uint8_t*>(&v);
                for (int k = 0; k < 16; ++k)
                    if (ga + k >= in && ga + k < frames_end) vb[k] = __ldg(ga + k);
            }
            *reinterpret_cast<uint4*>(s_rows + r * pitch + 16 * c) = v;
        }
    }
    __syncthreads();
    for (int xx = threadIdx.x; xx < nw; xx += blockDim.x) {
        const int xmin = __ldg(bounds + 2 * xx), xn = __ldg(bounds + 2 * xx + 1);
        int acc[kPilRows][3];
#pragma unroll
        for (int r = 0; r < kPilRows; ++r) acc[r][0] = acc[r][1] = acc[r][2] = 1 << (kPilBits - 1);
        const uint8_t* p0 = s_rows + ph[0] + xmin * 3;
        const uint8_t* p1 = s_rows + pitch + ph[kPilRows - 1] + xmin * 3;
        for (int x = 0; x < xn; ++x) {
            const int c = __ldg(kk_t + (size_t)x * nw + xx);
            acc[0][0] += (int)p0[3 * x] * c; acc[0][1] += (int)p0[3 * x + 1] * c; acc[0][2] += (int)p0[3 * x + 2] * c;
            acc[1][0] += (int)p1[3 * x] * c; acc[1][1] += (int)p1[3 * x + 1] * c; acc[1][2] += (int)p1[3 * x + 2] * c;
        }
#pragma unroll
        for (int r = 0; r < kPilRows; ++r) {
            if (r >= nr) continue;
            uint8_t* o = tmp + (((size_t)f * rows + r0 + r) * nw + xx) * 3;
            o[0] = (uint8_t)pil_clip8(acc[r][0]); o[1] = (uint8_t)pil_clip8(acc[r][1]); o[2] = (uint8_t)pil_clip8(acc[r][2]);
        }
    }
}

// pad canvas: every output element = pad_value (the content rectangle is overwritten by the vertical pass)
template <typename T>
__global__ void __launch_bounds__(256)
pil_fill_kernel(T* __restrict__ out, size_t total, int pad_value) {
    const T v = cvt<T>(pad_value);
    const size_t stride = (size_t)gridDim.x * blockDim.x;
    for (size_t t = (size_t)blockIdx.x * blockDim.x + threadIdx.x; t < total; t += stride) out[t] = v;
}

// vertical pass over the content rectangle + /255 + layout.  The vertical filter does not mix channels, so a
// row of the uint8 intermediate is a flat array of nw*3 bytes: one thread = 4 consecutive bytes, one 32-bit load
// per tap (adjacent threads read adjacent words), all loads of a tap window independent.
template <typename T>
__global__ void __launch_bounds__(256)
pil_vpass_kernel(const uint8_t* __restrict__ tmp, int n, int rows, int row0, int nw, int nh,
                 const int* __restrict__ bounds, const int* __restrict__ kk, int ksize, T* __restrict__ out,
                 int out_h, int out_w, int ox, int oy, int swap_rb, int nchw, int vertical) {
    const int row_bytes = nw * 3;
    const int groups = (row_bytes + 3) >> 2;
    const size_t total = (size_t)n * nh * groups;
    const size_t stride = (size_t)gridDim.x * blockDim.x;
    const bool vec = (row_bytes & 3) == 0 && (reinterpret_cast<uintptr_t>(tmp) & 3) == 0;
    const size_t plane = (size_t)out_h * out_w;
    for (size_t t = (size_t)blockIdx.x * blockDim.x + threadIdx.x; t < total; t += stride) {
        const int q = (int)(t % groups);
        const int yy = (int)((t / groups) % nh);
        const int f = (int)(t / ((size_t)groups * nh));
        const int j0 = q * 4, nb = min(4, row_bytes - j0);
        int v[4];
        if (vertical) {
            const int ymin = __ldg(bounds + 2 * yy) - row0, yn = __ldg(bounds + 2 * yy + 1);
            const int* k = kk + (size_t)yy * ksize;
            const uint8_t* src = tmp + ((size_t)f * rows + ymin) * row_bytes + j0;
            int a0 = 1 << (kPilBits - 1), a1 = a0, a2 = a0, a3 = a0;
            if (vec) {
                for (int r = 0; r < yn; ++r) {
                    const int c = __ldg(k + r);
                    const uint32_t w = __ldg(reinterpret_cast<const uint32_t*>(src + (size_t)r * row_bytes));
                    a0 += (int)(w & 0xffu) * c; a1 += (int)((w >> 8) & 0xffu) * c;
                    a2 += (int)((w >> 16) & 0xffu) * c; a3 += (int)(w >> 24) * c;
                }
            } else {
                for (int r = 0; r < yn; ++r) {
                    const int c = __ldg(k + r);
                    const uint8_t* qq = src + (size_t)r * row_bytes;
                    a0 += (int)__ldg(qq) * c;
                    if (nb > 1) a1 += (int)__ldg(qq + 1) * c;
                    if (nb > 2) a2 += (int)__ldg(qq + 2) * c;
                    if (nb > 3) a3 += (int)__ldg(qq + 3) * c;
                }
            }
            v[0] = pil_clip8(a0); v[1] = pil_clip8(a1); v[2] = pil_clip8(a2); v[3] = pil_clip8(a3);
        } else {
            const uint8_t* qq = tmp + ((size_t)f * rows + yy - row0) * row_bytes + j0;
            for (int b = 0; b < 4; ++b) v[b] = b < nb ? (int)__ldg(qq + b) : 0;
        }
        const int y = oy + yy;
#pragma unroll
        for (int b = 0; b < 4; ++b) {
            if (b >= nb) break;
            const int j = j0 + b, x = ox + j / 3;
            int c = j % 3;
            if (swap_rb) c = 2 - c;
            if (nchw) out[((size_t)f * 3 + c) * plane + (size_t)y * out_w + x] = cvt<T>(v[b]);
            else out[(((size_t)f * out_h + y) * out_w + x) * 3 + c] = cvt<T>(v[b]);
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
    // the tables depend on (W, nw, H, nh) only: a stream of equally sized frames builds and uploads them once
    hbp_pil_cache& pc = ctx->pil;
    if (!(pc.valid && pc.W == W && pc.nw == nw && pc.H == H && pc.nh == nh)) {
    pc.valid = false;
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
    {   // horizontal coefficients tap-major: kh_t[x * nw + xx]
        std::vector<int> kt(kh.size());
        for (int xx = 0; xx < nw; ++xx)
            for (int x = 0; x < ks_h; ++x) kt[(size_t)x * nw + xx] = kh[(size_t)xx * ks_h + x];
        kh.swap(kt);
    }
    tab.insert(tab.end(), bh.begin(), bh.end()); tab.insert(tab.end(), kh.begin(), kh.end());
    tab.insert(tab.end(), bv.begin(), bv.end()); tab.insert(tab.end(), kv.begin(), kv.end());
    HBP_CUDA(cudaMemcpyAsync(d_tab, tab.data(), n_tab * sizeof(int), cudaMemcpyHostToDevice, ctx->stream));
    HBP_CUDA(cudaStreamSynchronize(ctx->stream));             // `tab` is a local
    pc.W = W; pc.nw = nw; pc.H = H; pc.nh = nh; pc.ks_h = ks_h; pc.ks_v = ks_v; pc.row0 = row0; pc.rows = rows;
    pc.o_kh = bh.size(); pc.o_bv = pc.o_kh + kh.size(); pc.o_kv = pc.o_bv + bv.size();
    pc.valid = true;
    }
    const int ks_h = pc.ks_h, ks_v = pc.ks_v, row0 = pc.row0, rows = pc.rows;
    const int* d_tab = (const int*)ctx->scratch[SC_PRE_COEF];
    const int* d_bh = d_tab; const int* d_kh = d_tab + pc.o_kh;
    const int* d_bv = d_tab + pc.o_bv; const int* d_kv = d_tab + pc.o_kv;
    const uint8_t* tmp = in;
    int t_rows = H, t_row0 = 0;
    if (need_h) {
        uint8_t* d_tmp = (uint8_t*)hbp_scratch(ctx, SC_PRE_TMP, (size_t)n * rows * nw * 3);
        if (!d_tmp) return HBP_ERR_NOMEM;
        const int groups = (rows + kPilRows - 1) / kPilRows;
        const size_t smem = (size_t)kPilRows * (((size_t)W * 3 + 30) / 16 * 16);
        if (smem > 200 * 1024) { hbp_set_error("frame rows of %d pixels do not fit the horizontal pass", W); return HBP_ERR_INVALID; }
        if (!(ctx->attr_flags & ATTR_PIL)) {
            HBP_CUDA(cudaFuncSetAttribute(pil_hpass_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
            ctx->attr_flags |= ATTR_PIL;
        }
        pil_hpass_kernel<<<(unsigned)((size_t)n * groups), 256, smem, ctx->stream>>>(in, n, H, W, row0, rows, nw, d_bh, d_kh, ks_h, d_tmp);
        HBP_LAUNCH_CHECK(ctx);
        tmp = d_tmp; t_rows = rows; t_row0 = row0;
    }
    if (nw != out_w || nh != out_h) {
        const size_t total = (size_t)n * 3 * out_h * out_w;
        const int blocks = (int)std::min<size_t>((total + 255) / 256, (size_t)ctx->sm_count * 16);
        pil_fill_kernel<T><<<blocks, 256, 0, ctx->stream>>>((T*)out, total, pad_value);
        HBP_LAUNCH_CHECK(ctx);
    }
    const size_t total = (size_t)n * nh * (((size_t)nw * 3 + 3) / 4);
    const int blocks = (int)std::min<size_t>((total + 255) / 256, (size_t)ctx->sm_count * 32);
    pil_vpass_kernel<T><<<blocks, 256, 0, ctx->stream>>>(tmp, n, t_rows, t_row0, nw, nh, d_bv, d_kv, ks_v, (T*)out, out_h, out_w,
                                                         ox, oy, swap_rb, nchw, need_v ? 1 : 0);
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
