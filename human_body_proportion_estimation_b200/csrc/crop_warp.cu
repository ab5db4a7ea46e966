// K4: per-person crop = batched affine warp, cv2.warpAffine-exact bilinear,
// /255, NCHW fp16 (or fp32) HRNet input.
//
// Reference: models/conv.py:59-80 (tf.image.crop_and_resize of the /255 frame,
// NHWC->NCHW) and human_body_length_est/modules/pose_estimator.py:29-45.  The
// parity gate is cv2.warpAffine(INTER_LINEAR | WARP_INVERSE_MAP, BORDER_CONSTANT 0)
// (BASELINE.json north_star); its arithmetic (OpenCV imgwarp.cpp, restated in
// oracle/imgproc.py:warp_affine_cv2) is:
//     adelta[x] = rint(M00*x*1024)          bdelta[x] = rint(M10*x*1024)
//     X0[y] = rint((M01*y + M02)*1024) + 16  Y0[y] = rint((M11*y + M12)*1024) + 16
//     X = (X0[y] + adelta[x]) >> 5, Y likewise       (1/32 px)
//     sx = X >> 5, fx = X & 31 ; four taps, zero outside the frame;
//     weights (1-fy/32)(1-fx/32)... in float32.
// All intermediates of the blend are multiples of 2^-10 below 2^8, so float32
// evaluates them exactly in any association.
//
// Work decomposition: one CTA per (person, band of kBandRows output rows).
// The band's source bounding box is computed from the exact integer
// coordinates of its four corners (the map is monotone in x and in y), the
// u8 source rows of that box are staged in shared memory with 16-byte
// coalesced loads, and every thread then produces 8 consecutive output pixels
// per channel plane = one 16-byte fp16 store per plane.  Boxes whose band does
// not fit the shared-memory budget (extreme zoom-out / rotation) sample global
// memory directly through the read-only path.
// HBM-bound: source-box bytes + 3*out_h*out_w*2 bytes per person.
#include "hbp_internal.cuh"

namespace {

constexpr int kBandRows = 32;             // output rows per CTA: the per-CTA tables (column deltas, tap offsets, x weights) are built once per band
constexpr int kPassRows = 8;              // ... and the band is sampled in passes of <= 8 rows whose source rows fit the shared-memory budget
constexpr int kThreads = 256;
constexpr int kSmemBudget = 56 * 1024;    // per CTA: three CTAs of 256 threads per SM (80 registers per thread decide that, not this)
constexpr int kSmemPad = 16;              // bytes in front of / behind the staged box: a zero-weight tap may read up to 3 bytes outside it
constexpr int kMaxTabW = 512;             // output widths up to this use the per-CTA column tables

struct Affine {
    double m00, m01, m02, m10, m11, m12;
};

__device__ __forceinline__ void src_coord(const Affine& A, int x, int y, int& X, int& Y) {
    // separate roundings exactly as OpenCV (no FMA contraction)
    const int ad = __double2int_rn(__dmul_rn(__dmul_rn(A.m00, (double)x), 1024.0));
    const int bd = __double2int_rn(__dmul_rn(__dmul_rn(A.m10, (double)x), 1024.0));
    const int X0 = __double2int_rn(__dmul_rn(__dadd_rn(__dmul_rn(A.m01, (double)y), A.m02), 1024.0)) + 16;
    const int Y0 = __double2int_rn(__dmul_rn(__dadd_rn(__dmul_rn(A.m11, (double)y), A.m12), 1024.0)) + 16;
    X = (X0 + ad) >> 5;
    Y = (Y0 + bd) >> 5;
}

// uint8 -> float without the conversion unit (I2F runs on the quarter-rate XU pipe, the profile's
// busiest): 2^23 + b is exactly representable, so (float)b == as_float(0x4B000000 | b) - 2^23
__device__ __forceinline__ float u8_to_float(uint8_t b) { return __uint_as_float(0x4B000000u | (uint32_t)b) - 8388608.0f; }
// correctly rounded x / 255 without MUFU.RCP: one residual step on the product with the correctly
// rounded reciprocal (Markstein); bit-identical to __fdiv_rn(x, 255.0f) for the blend's value range
// (checked against it by tests/test_gpu_parity.py through the cv2-exact oracle)
__device__ __forceinline__ float div255(float x) {
    const float y = 1.0f / 255.0f;
    const float q = __fmul_rn(x, y);
    const float r = __fmaf_rn(-q, 255.0f, x);
    return __fmaf_rn(r, y, q);
}

template <typename OutT>
__device__ __forceinline__ OutT to_out(float v);
template <> __device__ __forceinline__ float to_out<float>(float v) { return v; }
template <> __device__ __forceinline__ __half to_out<__half>(float v) { return __float2half_rn(v); }

// Integer accumulator (sum of byte x weight in 1/1024 units, carried in the mantissa of 2^23: bits 0x4B000000 | acc)
// -> acc / 1024 / 255 in the output type.  fp32: exact int -> float, correctly rounded /255, exact 2^-10.
// fp16: ONE fused multiply-add, RN(acc * RN(1/261120)), gives the same fp16 value as rounding the correctly
// rounded fp32 quotient for every possible accumulator 0..261120 (exhaustive check: tests/test_host_cpu.py).
template <typename OutT>
__device__ __forceinline__ OutT finish_acc(unsigned bits);
template <> __device__ __forceinline__ float finish_acc<float>(unsigned bits) {
    return div255(__uint_as_float(bits) - 8388608.0f) * 0.0009765625f;
}
template <> __device__ __forceinline__ __half finish_acc<__half>(unsigned bits) {
    const float y = 1.0f / 261120.0f;
    return __float2half_rn(__fmaf_rn(__uint_as_float(bits), y, -8388608.0f * y));
}

template <typename OutT>
__global__ void __launch_bounds__(kThreads, 3)
crop_warp_kernel(const uint8_t* __restrict__ frames, int n_frames, int H, int W,
                 const double* __restrict__ Ms, const int* __restrict__ frame_idx, int P,
                 int out_h, int out_w, int swap_rb, OutT* __restrict__ out, const int* __restrict__ live) {
    extern __shared__ __align__(16) uint8_t smem_all[];
    if (live && (int)blockIdx.y >= *live) return;       // chained pipeline: person slots beyond the device-side count
    uint8_t* const smem = smem_all + kSmemPad;
    __shared__ int s_ad[kMaxTabW], s_bd[kMaxTabW];      // OpenCV's adelta / bdelta tables (per output column)
    // axis-aligned maps (every crop_and_resize box): per column the staged byte offset of the left tap and the
    // packed x-weights, rebuilt per pass (they depend on the pass's box)
    __shared__ __align__(16) int s_cx3[kMaxTabW];
    __shared__ __align__(16) unsigned s_wx[kMaxTabW];
    __shared__ int s_srow[2 * kBandRows];
    __shared__ int s_box[4];        // sx_min, sy_min, cols, rows (clamped to the frame)
    __shared__ int s_use_smem;

    const int p = blockIdx.y;
    const int row0 = blockIdx.x * kBandRows;
    const int rows = min(kBandRows, out_h - row0);
    Affine A;
    {
        const double* m = Ms + (size_t)p * 6;
        A.m00 = m[0]; A.m01 = m[1]; A.m02 = m[2]; A.m10 = m[3]; A.m11 = m[4]; A.m12 = m[5];
    }
    const bool tab = out_w <= kMaxTabW;
    const bool axis = tab && A.m01 == 0.0 && A.m10 == 0.0;
    if (tab)
        for (int x = threadIdx.x; x < out_w; x += kThreads) {
            s_ad[x] = __double2int_rn(__dmul_rn(__dmul_rn(A.m00, (double)x), 1024.0));
            s_bd[x] = axis ? 0 : __double2int_rn(__dmul_rn(__dmul_rn(A.m10, (double)x), 1024.0));      // (rint(0 * x * 1024) = 0)
        }
    int f = frame_idx[p];
    f = f < 0 ? 0 : (f >= n_frames ? n_frames - 1 : f);
    const uint8_t* __restrict__ src = frames + (size_t)f * H * W * 3;

    // The band is processed in passes of `rp` output rows (8, 4, 2 or 1): the largest whose source
    // box fits the shared-memory budget.  A small budget keeps many CTAs resident per SM (the
    // kernel is latency-bound: box -> stage -> sample are dependent phases), and big boxes
    // (4K frames, persons filling the frame) no longer fall back to sampling global memory.
    __shared__ int s_rp;
    auto box_of = [&](int r0, int nr, int* box, long long* need) {
        int xmin = INT_MAX, xmax = INT_MIN, ymin = INT_MAX, ymax = INT_MIN;
        const int cx[2] = {0, out_w - 1}, cy[2] = {r0, r0 + nr - 1};
#pragma unroll
        for (int a = 0; a < 2; ++a)
#pragma unroll
            for (int b = 0; b < 2; ++b) {
                int X, Y;
                src_coord(A, cx[a], cy[b], X, Y);
                xmin = min(xmin, X >> 5); xmax = max(xmax, X >> 5);
                ymin = min(ymin, Y >> 5); ymax = max(ymax, Y >> 5);
            }
        xmax += 1; ymax += 1;                       // right / bottom taps
        xmin = max(xmin, 0); ymin = max(ymin, 0);
        xmax = min(xmax, W - 1); ymax = min(ymax, H - 1);
        const int cols = xmax - xmin + 1, nrows = ymax - ymin + 1;
        box[0] = xmin; box[1] = ymin; box[2] = cols; box[3] = nrows;
        // axis-aligned maps stage only the two source rows every output row taps when that is fewer rows
        // than the whole span (down-scaling by more than 2 vertically)
        const int srows = axis ? min(nrows, 2 * nr) : nrows;
        *need = (cols > 0 && nrows > 0) ? (long long)srows * (((long long)cols * 3 + 15 + 15) / 16 * 16) : 0;
    };
    const size_t plane = (size_t)out_h * out_w;
    OutT* __restrict__ obase = out + (size_t)p * 3 * plane;
    __shared__ int s_xr[2];
    if (axis) {
        // axis-aligned: the source column range is the same for every output row -- two coordinate evaluations
        if (threadIdx.x == 0) {
            int Xa, Xb, Yd;
            src_coord(A, 0, row0, Xa, Yd);
            src_coord(A, out_w - 1, row0, Xb, Yd);
            int xmin = min(Xa >> 5, Xb >> 5), xmax = max(Xa >> 5, Xb >> 5) + 1;
            xmin = max(xmin, 0); xmax = min(xmax, W - 1);
            s_xr[0] = xmin; s_xr[1] = xmax - xmin + 1;
        }
        __syncthreads();
    // ---- row mode (every axis-aligned map whose source rows are <= 1536 bytes wide: all crop_and_resize / resize boxes up
    // to ~500 source pixels): after the per-band tables NO block barrier is left.  Every warp owns output rows ry = warp,
    // warp + 8, ...: it stages the two source rows the row taps in its private shared-memory slice (16-byte coalesced loads,
    // the next row's loads already in flight while the current row is sampled), samples its 2 x 3 pixels per lane with the
    // integer blend below, and moves on.  Same arithmetic as the band path, different schedule.
    {
        const int bx0 = s_xr[0], bcols = s_xr[1];
        const int pitch = ((bcols * 3 + 15 + 15) / 16) * 16;
        const int chunks_per_row = pitch / 16;
        constexpr int kWarpsC = kThreads / 32;
        if (bcols > 0 && kWarpsC * 2 * (pitch + 16) <= kSmemBudget + kSmemPad && (out_w & 1) == 0) {
            const int lane = threadIdx.x & 31, wrp = threadIdx.x >> 5;
            {
                const int X0c = __double2int_rn(__dmul_rn(A.m02, 1024.0)) + 16;
                for (int x = threadIdx.x; x < ((out_w + 7) & ~7); x += kThreads) {
                    const int X = (X0c + s_ad[min(x, out_w - 1)]) >> 5;
                    const int sx = X >> 5, fx = X & 31;
                    const unsigned wx0 = (unsigned)sx < (unsigned)W ? 32u - fx : 0u, wx1 = (unsigned)(sx + 1) < (unsigned)W ? (unsigned)fx : 0u;
                    s_cx3[x] = (min(max(sx, bx0 - 1), bx0 + bcols - 1) - bx0) * 3;
                    s_wx[x] = wx0 | (wx1 << 16);
                }
            }
            __syncthreads();
            uint8_t* const wbuf = smem_all + wrp * 2 * (pitch + 16);        // [tap row][16-byte front pad | pitch]
            const uint8_t* frame_end = frames + (size_t)n_frames * H * W * 3;
            const int src_lo = (int)(reinterpret_cast<uintptr_t>(src) & 15);
            const bool pair_ok = (plane & 1) == 0 && (reinterpret_cast<uintptr_t>(obase) & (2 * sizeof(OutT) - 1)) == 0;
            uint4 v[2][3];
            int srow[2] = {0, 0};
            auto request = [&](int ry) {            // global loads of the two source rows of output row row0 + ry -> registers
                const int y = row0 + ry;
                const int Y0 = __double2int_rn(__dmul_rn(__dadd_rn(__dmul_rn(A.m11, (double)y), A.m12), 1024.0)) + 16;
                const int sy = (Y0 >> 5) >> 5;
#pragma unroll
                for (int t = 0; t < 2; ++t) {
                    const int r = min(max(sy + t, 0), H - 1);       // (a tap outside the frame has weight 0: any valid row will do)
                    srow[t] = r;
                    const uint8_t* g0 = src + ((size_t)r * W + bx0) * 3;
                    const uint8_t* ga = reinterpret_cast<const uint8_t*>(reinterpret_cast<uintptr_t>(g0) & ~uintptr_t(15));
                    const bool inside = ga >= frames && ga + (size_t)16 * chunks_per_row <= frame_end;
#pragma unroll
                    for (int k = 0; k < 3; ++k) {
                        const int c = lane + 32 * k;
                        v[t][k] = make_uint4(0, 0, 0, 0);
                        if (c >= chunks_per_row) continue;
                        if (inside) {
                            v[t][k] = __ldg(reinterpret_cast<const uint4*>(ga) + c);
                        } else {                                    // first / last bytes of the allocation
                            uint8_t* vb = reinterpret_cast<uint8_t*>(&v[t][k]);
                            for (int b = 0; b < 16; ++b)
                                if (ga + 16 * c + b >= frames && ga + 16 * c + b < frame_end) vb[b] = __ldg(ga + 16 * c + b);
                        }
                    }
                }
            };
            if (wrp < rows) request(wrp);
            for (int ry = wrp; ry < rows; ry += kWarpsC) {
                const int y = row0 + ry;
                const int Y0 = __double2int_rn(__dmul_rn(__dadd_rn(__dmul_rn(A.m11, (double)y), A.m12), 1024.0)) + 16;
                const int Y = Y0 >> 5, sy = Y >> 5, fy = Y & 31;
                const unsigned wy0 = (unsigned)sy < (unsigned)H ? 32u - fy : 0u, wy1 = (unsigned)(sy + 1) < (unsigned)H ? (unsigned)fy : 0u;
                const int base0 = 16 + ((src_lo + (srow[0] * W + bx0) * 3) & 15);
                const int base1 = (pitch + 16) + 16 + ((src_lo + (srow[1] * W + bx0) * 3) & 15);
#pragma unroll
                for (int t = 0; t < 2; ++t)
#pragma unroll
                    for (int k = 0; k < 3; ++k)
                        if (lane + 32 * k < chunks_per_row) *reinterpret_cast<uint4*>(wbuf + t * (pitch + 16) + 16 + 16 * (lane + 32 * k)) = v[t][k];
                if (chunks_per_row > 96) {
                    // rows wider than 1536 bytes (source boxes beyond ~500 pixels): the rest of the row, not prefetched
#pragma unroll
                    for (int t = 0; t < 2; ++t) {
                        const uint8_t* g0 = src + ((size_t)srow[t] * W + bx0) * 3;
                        const uint8_t* ga = reinterpret_cast<const uint8_t*>(reinterpret_cast<uintptr_t>(g0) & ~uintptr_t(15));
                        const bool inside = ga >= frames && ga + (size_t)16 * chunks_per_row <= frame_end;
                        for (int c = lane + 96; c < chunks_per_row; c += 32) {
                            uint4 q = make_uint4(0, 0, 0, 0);
                            if (inside) q = __ldg(reinterpret_cast<const uint4*>(ga) + c);
                            else {
                                uint8_t* vb = reinterpret_cast<uint8_t*>(&q);
                                for (int b = 0; b < 16; ++b)
                                    if (ga + 16 * c + b >= frames && ga + 16 * c + b < frame_end) vb[b] = __ldg(ga + 16 * c + b);
                            }
                            *reinterpret_cast<uint4*>(wbuf + t * (pitch + 16) + 16 + 16 * c) = q;
                        }
                    }
                }
                __syncwarp();
                if (ry + kWarpsC < rows) request(ry + kWarpsC);       // next row's loads fly while this one is sampled
                OutT* orow0 = obase + (size_t)y * out_w + (swap_rb ? 2 * plane : 0);
                OutT* orow1 = obase + (size_t)y * out_w + plane;
                OutT* orow2 = obase + (size_t)y * out_w + (swap_rb ? 0 : 2 * plane);
                for (int x0 = 2 * lane; x0 < out_w; x0 += 64) {
                    const int2 c3 = *reinterpret_cast<const int2*>(&s_cx3[x0]);
                    const uint2 wx = *reinterpret_cast<const uint2*>(&s_wx[x0]);
                    OutT r[3][2];
#pragma unroll
                    for (int k = 0; k < 2; ++k) {
                        const int cc = k ? c3.y : c3.x;
                        const unsigned wxx = k ? wx.y : wx.x;
                        const unsigned wr0 = wxx * wy0, wr1 = wxx * wy1;
                        unsigned acc0 = 0x4B000000u, acc1 = 0x4B000000u, acc2 = 0x4B000000u;
                        auto row = [&](int b, unsigned wr) {
                            const uint32_t* wp = reinterpret_cast<const uint32_t*>(wbuf + (b & ~3));
                            const uint32_t lo = wp[0], mid = wp[1], hi = wp[2];
                            const unsigned sh = (unsigned)b << 3;                  // funnel shift takes the amount mod 32
                            const uint32_t v0 = __funnelshift_r(lo, mid, sh), v1 = __funnelshift_r(mid, hi, sh);
                            acc0 = __dp2a_lo(wr, __byte_perm(v0, v1, 0x4430), acc0);
                            acc1 = __dp2a_lo(wr, __byte_perm(v0, v1, 0x4441), acc1);
                            acc2 = __dp2a_lo(wr, __byte_perm(v0, v1, 0x4452), acc2);
                        };
                        row(base0 + cc, wr0);
                        row(base1 + cc, wr1);
                        r[0][k] = finish_acc<OutT>(acc0);
                        r[1][k] = finish_acc<OutT>(acc1);
                        r[2][k] = finish_acc<OutT>(acc2);
                    }
                    OutT* const op[3] = {orow0 + x0, orow1 + x0, orow2 + x0};
                    if (pair_ok) {
#pragma unroll
                        for (int c = 0; c < 3; ++c) {
                            if (sizeof(OutT) == 2) *reinterpret_cast<uint32_t*>(op[c]) = *reinterpret_cast<const uint32_t*>(r[c]);
                            else *reinterpret_cast<uint2*>(op[c]) = *reinterpret_cast<const uint2*>(r[c]);
                        }
                    } else {
#pragma unroll
                        for (int c = 0; c < 3; ++c) { op[c][0] = r[c][0]; op[c][1] = r[c][1]; }
                    }
                }
                __syncwarp();                   // the slice is restaged for this warp's next row
            }
            return;
        }
    }
    }
    // (all box arithmetic is double precision, slow on this part: spread it over threads instead of leaving 255 threads
    // at a barrier behind thread 0 -- the round-1 profile had stall_barrier on top)
    __shared__ int s_fit[4];
    __shared__ int s_pbox[kBandRows][5];    // per pass: x0, y0, cols, rows, staged?
    if (threadIdx.x < 4) {
        const int rp_t = kPassRows >> threadIdx.x;      // 8, 4, 2, 1
        int box[4];
        long long need;
        box_of(row0, min(rp_t, rows), box, &need);
        s_fit[threadIdx.x] = need <= kSmemBudget ? 1 : 0;
    }
    __syncthreads();
    if (threadIdx.x == 0) s_rp = s_fit[0] ? 8 : s_fit[1] ? 4 : s_fit[2] ? 2 : 1;
    __syncthreads();
    const int rp = s_rp;
    {
        const int n_pass = (rows + rp - 1) / rp;
        if ((int)threadIdx.x < n_pass) {
            const int pr0_t = row0 + (int)threadIdx.x * rp;
            int box[4];
            long long need;
            box_of(pr0_t, min(rp, row0 + rows - pr0_t), box, &need);
            s_pbox[threadIdx.x][0] = box[0]; s_pbox[threadIdx.x][1] = box[1]; s_pbox[threadIdx.x][2] = box[2]; s_pbox[threadIdx.x][3] = box[3];
            s_pbox[threadIdx.x][4] = (box[2] > 0 && box[3] > 0 && need <= kSmemBudget) ? 1 : 0;
        }
    }
    __syncthreads();
    const int groups = (out_w + 7) / 8;                 // 8 output pixels per thread-iteration
    bool tables_built = false;
    for (int pr0 = row0; pr0 < row0 + rows; pr0 += rp) {
    const int prow = min(rp, row0 + rows - pr0);
    const int* pb = s_pbox[(pr0 - row0) / rp];
    const int bx0 = pb[0], by0 = pb[1], bcols = pb[2], brows = pb[3];
    const bool staged = pb[4] != 0;
    const int pitch = ((bcols * 3 + 15 + 15) / 16) * 16;   // bytes per staged row

    const bool slots = staged && axis && brows > 2 * prow;      // staged row 2*ry + t = source row of tap t of output row pr0 + ry
    const int srows = slots ? 2 * prow : brows;
    if (slots) {
        if (threadIdx.x < 2 * prow) {
            const int y = pr0 + ((int)threadIdx.x >> 1);
            const int Y0 = __double2int_rn(__dmul_rn(__dadd_rn(__dmul_rn(A.m11, (double)y), A.m12), 1024.0)) + 16;
            const int sy = (Y0 >> 5) >> 5;
            s_srow[threadIdx.x] = min(max(sy + ((int)threadIdx.x & 1), by0), by0 + brows - 1);
        }
        __syncthreads();
    }
    if (staged && axis && !tables_built) {
        // (an axis-aligned map has the same source column range in every row: one table per band)
        tables_built = true;
        const int X0c = __double2int_rn(__dmul_rn(A.m02, 1024.0)) + 16;
        for (int x = threadIdx.x; x < ((out_w + 7) & ~7); x += kThreads) {
            const int X = (X0c + s_ad[min(x, out_w - 1)]) >> 5;
            const int sx = X >> 5, fx = X & 31;
            const unsigned wx0 = (unsigned)sx < (unsigned)W ? 32u - fx : 0u, wx1 = (unsigned)(sx + 1) < (unsigned)W ? (unsigned)fx : 0u;
            s_cx3[x] = (min(max(sx, bx0 - 1), bx0 + bcols - 1) - bx0) * 3;
            s_wx[x] = wx0 | (wx1 << 16);
        }
    }
    if (staged) {
        // Stage rows [by0, by0+brows) x byte range of columns [bx0, bx0+bcols).
        // Row r lives at smem[r*pitch + (addr & 15) ...]: the copy is done in
        // 16-byte units aligned in GLOBAL memory, so every row keeps its own
        // sub-16 phase `ph`; taps index with that phase.
        const int chunks_per_row = pitch / 16;
        const uint8_t* frame_end = frames + (size_t)n_frames * H * W * 3;
        // One warp per staged row, lanes over its 16-byte chunks: the row's base pointer, its alignment and the bounds test
        // are computed once per row and warp (the flat chunk loop of round 1 spent ~40 instructions per chunk on a division,
        // two table reads and 64-bit address arithmetic: 45 % of the kernel's instructions, profiles/r02_crop_ncu.md).
        // Two rows' loads (up to six 16-byte loads per lane) are in flight per warp before the first shared-memory store.
        const int lane_s = threadIdx.x & 31, wrp_s = threadIdx.x >> 5;
        constexpr int kWarpsC = kThreads / 32;
        for (int rb = wrp_s; rb < srows; rb += 2 * kWarpsC) {
            uint4 v[2][3];
            bool on[2];
            int nch[2];
#pragma unroll
            for (int u = 0; u < 2; ++u) {
                const int r = rb + u * kWarpsC;
                on[u] = r < srows;
                nch[u] = 0;
                if (!on[u]) continue;
                const uint8_t* g0 = src + ((size_t)(slots ? s_srow[r] : by0 + r) * W + bx0) * 3;
                const uint8_t* ga = reinterpret_cast<const uint8_t*>(reinterpret_cast<uintptr_t>(g0) & ~uintptr_t(15));
                const bool inside = ga >= frames && ga + (size_t)16 * chunks_per_row <= frame_end;
#pragma unroll
                for (int k = 0; k < 3; ++k) {
                    const int c = lane_s + 32 * k;
                    v[u][k] = make_uint4(0, 0, 0, 0);
                    if (c >= chunks_per_row) continue;
                    nch[u] = k + 1;
                    if (inside) {
                        v[u][k] = __ldg(reinterpret_cast<const uint4*>(ga) + c);
                    } else {                                    // first / last bytes of the allocation
                        uint8_t* vb = reinterpret_cast<uint8_t*>(&v[u][k]);
                        for (int b = 0; b < 16; ++b)
                            if (ga + 16 * c + b >= frames && ga + 16 * c + b < frame_end) vb[b] = __ldg(ga + 16 * c + b);
                    }
                }
                // (rows wider than 96 chunks = 1536 bytes: the remaining chunks one at a time)
                for (int c = lane_s + 96; c < chunks_per_row; c += 32) {
                    uint4 q = make_uint4(0, 0, 0, 0);
                    if (inside) q = __ldg(reinterpret_cast<const uint4*>(ga) + c);
                    else {
                        uint8_t* vb = reinterpret_cast<uint8_t*>(&q);
                        for (int b = 0; b < 16; ++b)
                            if (ga + 16 * c + b >= frames && ga + 16 * c + b < frame_end) vb[b] = __ldg(ga + 16 * c + b);
                    }
                    *reinterpret_cast<uint4*>(smem + r * pitch + 16 * c) = q;
                }
            }
#pragma unroll
            for (int u = 0; u < 2; ++u) {
                if (!on[u]) continue;
                const int r = rb + u * kWarpsC;
#pragma unroll
                for (int k = 0; k < 3; ++k)
                    if (k < nch[u]) *reinterpret_cast<uint4*>(smem + r * pitch + 16 * (lane_s + 32 * k)) = v[u][k];
            }
        }
        __syncthreads();
    }

    if (staged && axis) {
        // axis-aligned: one warp per output row, a lane takes pixel pairs x = 2*lane + 64*j.  The row terms
        // (sy, fy, staged row bases) are per warp-row, the column terms come from the per-pass tables, and
        // neighbouring lanes read neighbouring staged bytes (the 8-pixels-per-thread layout of the general
        // path puts lanes 8 pixels apart: 4-way shared-memory bank conflicts).  ~35 instructions per pixel.
        const int src_lo = (int)(reinterpret_cast<uintptr_t>(src) & 15);
        const int lane = threadIdx.x & 31, wrp = threadIdx.x >> 5;
        for (int ry = wrp; ry < prow; ry += kThreads / 32) {
            const int y = pr0 + ry;
            const int Y0 = __double2int_rn(__dmul_rn(__dadd_rn(__dmul_rn(A.m11, (double)y), A.m12), 1024.0)) + 16;
            const int Y = Y0 >> 5, sy = Y >> 5, fy = Y & 31;
            const unsigned wy0 = (unsigned)sy < (unsigned)H ? 32u - fy : 0u, wy1 = (unsigned)(sy + 1) < (unsigned)H ? (unsigned)fy : 0u;
            const int r0 = min(max(sy, by0), by0 + brows - 1), r1 = min(max(sy + 1, by0), by0 + brows - 1);
            const int base0 = (slots ? 2 * ry : r0 - by0) * pitch + ((src_lo + (r0 * W + bx0) * 3) & 15);
            const int base1 = (slots ? 2 * ry + 1 : r1 - by0) * pitch + ((src_lo + (r1 * W + bx0) * 3) & 15);
            // channel c goes to plane (swap_rb ? 2 - c : c): the swap is a choice of destination, not of value
            OutT* orow0 = obase + (size_t)y * out_w + (swap_rb ? 2 * plane : 0);
            OutT* orow1 = obase + (size_t)y * out_w + plane;
            OutT* orow2 = obase + (size_t)y * out_w + (swap_rb ? 0 : 2 * plane);
            // pair stores need every row start 2-element aligned: one uniform test per kernel, not one per store
            const bool pair_ok = (out_w & 1) == 0 && (plane & 1) == 0 && (reinterpret_cast<uintptr_t>(obase) & (2 * sizeof(OutT) - 1)) == 0;
            for (int x0 = 2 * lane; x0 < out_w; x0 += 64) {
                const int2 c3 = *reinterpret_cast<const int2*>(&s_cx3[x0]);
                const uint2 wx = *reinterpret_cast<const uint2*>(&s_wx[x0]);
                OutT r[3][2];
#pragma unroll
                for (int k = 0; k < 2; ++k) {
                    const int cc = k ? c3.y : c3.x;
                    const unsigned wxx = k ? wx.y : wx.x;
                    const unsigned wr0 = wxx * wy0, wr1 = wxx * wy1;
                    unsigned acc0 = 0x4B000000u, acc1 = 0x4B000000u, acc2 = 0x4B000000u;
                    auto row = [&](int b, unsigned wr) {
                        const uint32_t* wp = reinterpret_cast<const uint32_t*>(smem + (b & ~3));
                        const uint32_t lo = wp[0], mid = wp[1], hi = wp[2];
                        const unsigned sh = (unsigned)b << 3;                  // funnel shift takes the amount mod 32
                        const uint32_t v0 = __funnelshift_r(lo, mid, sh), v1 = __funnelshift_r(mid, hi, sh);
                        acc0 = __dp2a_lo(wr, __byte_perm(v0, v1, 0x4430), acc0);
                        acc1 = __dp2a_lo(wr, __byte_perm(v0, v1, 0x4441), acc1);
                        acc2 = __dp2a_lo(wr, __byte_perm(v0, v1, 0x4452), acc2);
                    };
                    row(base0 + cc, wr0);
                    row(base1 + cc, wr1);
                    r[0][k] = finish_acc<OutT>(acc0);
                    r[1][k] = finish_acc<OutT>(acc1);
                    r[2][k] = finish_acc<OutT>(acc2);
                }
                OutT* const op[3] = {orow0 + x0, orow1 + x0, orow2 + x0};
                if (pair_ok) {                  // x0 even and out_w even: the pair is inside the row
#pragma unroll
                    for (int c = 0; c < 3; ++c) {
                        if (sizeof(OutT) == 2) *reinterpret_cast<uint32_t*>(op[c]) = *reinterpret_cast<const uint32_t*>(r[c]);
                        else *reinterpret_cast<uint2*>(op[c]) = *reinterpret_cast<const uint2*>(r[c]);
                    }
                } else {
#pragma unroll
                    for (int c = 0; c < 3; ++c) {
                        op[c][0] = r[c][0];
                        if (x0 + 1 < out_w) op[c][1] = r[c][1];
                    }
                }
            }
        }
    } else
    for (int t = threadIdx.x; t < prow * groups; t += kThreads) {
        const int ry = t / groups, gx = t - ry * groups;
        const int y = pr0 + ry, xbeg = gx * 8;
        OutT res[3][8];
        // per-row constants
        const int X0 = __double2int_rn(__dmul_rn(__dadd_rn(__dmul_rn(A.m01, (double)y), A.m02), 1024.0)) + 16;
        const int Y0 = __double2int_rn(__dmul_rn(__dadd_rn(__dmul_rn(A.m11, (double)y), A.m12), 1024.0)) + 16;
        if (staged && tab) {
            // Integer blend: the bilinear weights are (32-fy|fy) x (32-fx|fx) in 1/1024 units and the taps are
            // bytes, so sum(b*w) <= 255*1024 is exact in int32 and equals 1024 x the float32 sum of the
            // reference arithmetic (which is exact as well).  Out-of-frame taps get weight 0 and a clamped
            // address.  The two x-taps of a row are 6 consecutive bytes: three aligned 32-bit loads, two
            // funnel shifts, one PRMT + one DP2A per channel and row.
            const int src_lo = (int)(reinterpret_cast<uintptr_t>(src) & 15);
#pragma unroll
            for (int k = 0; k < 8; ++k) {
                const int x = min(xbeg + k, out_w - 1);
                const int X = (X0 + s_ad[x]) >> 5, Y = (Y0 + s_bd[x]) >> 5;
                const int sx = X >> 5, sy = Y >> 5;
                const int fx = X & 31, fy = Y & 31;
                const unsigned wx0 = (unsigned)sx < (unsigned)W ? 32u - fx : 0u, wx1 = (unsigned)(sx + 1) < (unsigned)W ? (unsigned)fx : 0u;
                const unsigned wy0 = (unsigned)sy < (unsigned)H ? 32u - fy : 0u, wy1 = (unsigned)(sy + 1) < (unsigned)H ? (unsigned)fy : 0u;
                const unsigned wx = wx0 | (wx1 << 16);                   // packed pair; products below stay < 2^16 per half
                const unsigned wr0 = wx * wy0, wr1 = wx * wy1;
                const int cx = min(max(sx, bx0 - 1), bx0 + bcols - 1) - bx0;          // -1 .. bcols-1 (the pad covers -1 and the right overhang)
                const int r0 = min(max(sy, by0), by0 + brows - 1), r1 = min(max(sy + 1, by0), by0 + brows - 1);
                unsigned acc[3] = {0x4B000000u, 0x4B000000u, 0x4B000000u};
                auto row = [&](int yy, unsigned wr) {
                    const int ph = (src_lo + (yy * W + bx0) * 3) & 15;
                    const int b = (yy - by0) * pitch + ph + cx * 3;
                    const uint32_t* wp = reinterpret_cast<const uint32_t*>(smem + (b & ~3));
                    const uint32_t lo = wp[0], mid = wp[1], hi = wp[2];
                    const unsigned sh = (unsigned)(b & 3) * 8u;
                    const uint32_t v0 = __funnelshift_r(lo, mid, sh), v1 = __funnelshift_r(mid, hi, sh);
                    acc[0] = __dp2a_lo(wr, __byte_perm(v0, v1, 0x4430), acc[0]);
                    acc[1] = __dp2a_lo(wr, __byte_perm(v0, v1, 0x4441), acc[1]);
                    acc[2] = __dp2a_lo(wr, __byte_perm(v0, v1, 0x4452), acc[2]);
                };
                row(r0, wr0);
                row(r1, wr1);
                res[0][k] = finish_acc<OutT>(swap_rb ? acc[2] : acc[0]);
                res[1][k] = finish_acc<OutT>(acc[1]);
                res[2][k] = finish_acc<OutT>(swap_rb ? acc[0] : acc[2]);
            }
        } else
#pragma unroll
        for (int k = 0; k < 8; ++k) {
            const int x = xbeg + k;
            const int ad = __double2int_rn(__dmul_rn(__dmul_rn(A.m00, (double)x), 1024.0));
            const int bd = __double2int_rn(__dmul_rn(__dmul_rn(A.m10, (double)x), 1024.0));
            const int X = (X0 + ad) >> 5, Y = (Y0 + bd) >> 5;
            const int sx = X >> 5, sy = Y >> 5;
            const float fx = (float)(X & 31) * 0.03125f, fy = (float)(Y & 31) * 0.03125f;
            const float w00 = (1.f - fy) * (1.f - fx), w01 = (1.f - fy) * fx;
            const float w10 = fy * (1.f - fx), w11 = fy * fx;
            const bool in_x0 = sx >= 0 && sx < W, in_x1 = sx + 1 >= 0 && sx + 1 < W;
            const bool in_y0 = sy >= 0 && sy < H, in_y1 = sy + 1 >= 0 && sy + 1 < H;
            float acc[3] = {0.f, 0.f, 0.f};
            auto tap = [&](int yy, int xx, bool ok, float wgt) {
                if (!ok) return;
                const uint8_t* q;
                if (staged) {
                    const uint8_t* g0 = src + ((size_t)yy * W + bx0) * 3;
                    const int ph = (int)(reinterpret_cast<uintptr_t>(g0) & 15);
                    q = smem + (size_t)(yy - by0) * pitch + ph + (xx - bx0) * 3;
                    acc[0] += u8_to_float(q[0]) * wgt; acc[1] += u8_to_float(q[1]) * wgt; acc[2] += u8_to_float(q[2]) * wgt;
                } else {
                    q = src + ((size_t)yy * W + xx) * 3;
                    acc[0] += u8_to_float(__ldg(q)) * wgt; acc[1] += u8_to_float(__ldg(q + 1)) * wgt; acc[2] += u8_to_float(__ldg(q + 2)) * wgt;
                }
            };
            tap(sy, sx, in_y0 && in_x0, w00);
            tap(sy, sx + 1, in_y0 && in_x1, w01);
            tap(sy + 1, sx, in_y1 && in_x0, w10);
            tap(sy + 1, sx + 1, in_y1 && in_x1, w11);
            res[0][k] = to_out<OutT>(div255(swap_rb ? acc[2] : acc[0]));
            res[1][k] = to_out<OutT>(div255(acc[1]));
            res[2][k] = to_out<OutT>(div255(swap_rb ? acc[0] : acc[2]));
        }
#pragma unroll
        for (int c = 0; c < 3; ++c) {
            OutT* o = obase + c * plane + (size_t)y * out_w + xbeg;
            if (xbeg + 8 <= out_w && (reinterpret_cast<uintptr_t>(o) & 15) == 0 && sizeof(OutT) == 2) {
                *reinterpret_cast<uint4*>(o) = *reinterpret_cast<const uint4*>(res[c]);
            } else if (xbeg + 8 <= out_w && (reinterpret_cast<uintptr_t>(o) & 15) == 0 && sizeof(OutT) == 4) {
                reinterpret_cast<uint4*>(o)[0] = reinterpret_cast<const uint4*>(res[c])[0];
                reinterpret_cast<uint4*>(o)[1] = reinterpret_cast<const uint4*>(res[c])[1];
            } else {
                for (int k = 0; k < 8 && xbeg + k < out_w; ++k) o[k] = res[c][k];
            }
        }
    }
    __syncthreads();            // the next pass overwrites the staged rows
    }
}

}  // namespace

int k_crop_warp(hbp_ctx* ctx, const uint8_t* frames, int n_frames, int h, int w, const double* M,
                const int* frame_idx, int P, int out_h, int out_w, int swap_rb, void* out,
                int out_dtype, const int* live) {
    if (P <= 0) return HBP_OK;
    if (!(ctx->attr_flags & ATTR_CROP)) {
        HBP_CUDA(cudaFuncSetAttribute(crop_warp_kernel<__half>, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemBudget + 2 * kSmemPad));
        HBP_CUDA(cudaFuncSetAttribute(crop_warp_kernel<float>, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemBudget + 2 * kSmemPad));
        ctx->attr_flags |= ATTR_CROP;
    }
    dim3 grid((out_h + kBandRows - 1) / kBandRows, P);      // (kBandRows output rows per CTA)
    if (out_dtype == HBP_F16)
        crop_warp_kernel<__half><<<grid, kThreads, kSmemBudget + 2 * kSmemPad, ctx->stream>>>(
            frames, n_frames, h, w, M, frame_idx, P, out_h, out_w, swap_rb, (__half*)out, live);
    else
        crop_warp_kernel<float><<<grid, kThreads, kSmemBudget + 2 * kSmemPad, ctx->stream>>>(
            frames, n_frames, h, w, M, frame_idx, P, out_h, out_w, swap_rb, (float*)out, live);
    HBP_LAUNCH_CHECK(ctx);
    return HBP_OK;
}
