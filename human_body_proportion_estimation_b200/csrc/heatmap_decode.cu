// K6: heatmap argmax decode fused with the box remap, per-joint gate and the
// 11 body-segment lengths.
//
// Reference semantics (paths relative to the reference's human_body_length_est/):
//   decode ........ modules/pose_estimator.py:74-99   (first-index argmax, >0 mask)
//   remap + gate .. person_det_pose_edet4_trtserver.py:145-168
//   lengths ....... modules/pose_estimator.py:130-200
//
// Layout: heatmaps (P,J,Hh,Wh) contiguous, fp32 or fp16.  One CTA per person,
// one warp per joint: each lane streams 16-byte vectors of its joint's map
// (coalesced 512 B per warp instruction), keeps a running (value,index) with
// numpy's ordering (NaN beats everything, ties -> lowest index) and the warp
// finishes with a shuffle reduction.  Warp 0 then does the geometry for the
// person from shared memory.  HBM-bound: J*Hh*Wh*e bytes read per person,
// J*12 + 11*4 + 12 bytes written.
//
// Every float operation that the reference performs with one IEEE rounding is
// written with the __f*_rn intrinsics so that no FMA contraction can change a
// bit.
#include "hbp_internal.cuh"

namespace {

constexpr int kMaxJ = 32;
constexpr int kWarps = 9;        // two joints per warp: small CTAs, seven resident per SM, so one CTA's geometry tail overlaps the others' streaming
constexpr int kThreads = kWarps * 32;

struct Best {
    float v;
    int i;
};

// numpy argmax order: a "beats" b when a is NaN and b is not, or a > b; equal
// (or both NaN) -> the lower index wins.
__device__ __forceinline__ bool beats(float av, int ai, float bv, int bi) {
    const bool an = av != av, bn = bv != bv;
    if (an || bn) return an && (!bn || ai < bi);
    return av > bv || (av == bv && ai < bi);
}

__device__ __forceinline__ void take(Best& b, float v, int i) {
    // within one lane indices only grow, so strict "greater" keeps the first
    if (v > b.v || (v != v && b.v == b.v)) { b.v = v; b.i = i; }
}

__device__ __forceinline__ Best warp_best(Best b) {
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) {
        float ov = __shfl_xor_sync(0xffffffffu, b.v, off);
        int oi = __shfl_xor_sync(0xffffffffu, b.i, off);
        if (beats(ov, oi, b.v, b.i)) { b.v = ov; b.i = oi; }
    }
    return b;
}

__device__ __forceinline__ Best scan_f32(const float* __restrict__ m, int n, int lane) {
    Best b{-INFINITY, 0x7fffffff};
    bool started = false;
    auto feed = [&](float v, int i) {
        if (!started) { b.v = v; b.i = i; started = true; } else take(b, v, i);
    };
    if ((reinterpret_cast<uintptr_t>(m) & 15) == 0) {
        const int n4 = n >> 2;
        const float4* m4 = reinterpret_cast<const float4*>(m);
        int i = lane;
        for (; i + 96 < n4; i += 128) {          // 4 independent 16 B loads in flight
            float4 a = __ldg(m4 + i), c = __ldg(m4 + i + 32), d = __ldg(m4 + i + 64), e = __ldg(m4 + i + 96);
            feed(a.x, 4 * i); feed(a.y, 4 * i + 1); feed(a.z, 4 * i + 2); feed(a.w, 4 * i + 3);
            feed(c.x, 4 * i + 128); feed(c.y, 4 * i + 129); feed(c.z, 4 * i + 130); feed(c.w, 4 * i + 131);
            feed(d.x, 4 * i + 256); feed(d.y, 4 * i + 257); feed(d.z, 4 * i + 258); feed(d.w, 4 * i + 259);
            feed(e.x, 4 * i + 384); feed(e.y, 4 * i + 385); feed(e.z, 4 * i + 386); feed(e.w, 4 * i + 387);
        }
        for (; i < n4; i += 32) {
            float4 a = __ldg(m4 + i);
            feed(a.x, 4 * i); feed(a.y, 4 * i + 1); feed(a.z, 4 * i + 2); feed(a.w, 4 * i + 3);
        }
        for (int t = (n4 << 2) + lane; t < n; t += 32) feed(__ldg(m + t), t);
    } else {
        for (int t = lane; t < n; t += 32) feed(__ldg(m + t), t);
    }
    if (!started) { b.v = -INFINITY; b.i = 0x7fffffff; }
    return b;
}

__device__ __forceinline__ Best scan_f16(const __half* __restrict__ m, int n, int lane) {
    Best b{-INFINITY, 0x7fffffff};
    bool started = false;
    auto feed = [&](float v, int i) {
        if (!started) { b.v = v; b.i = i; started = true; } else take(b, v, i);
    };
    if ((reinterpret_cast<uintptr_t>(m) & 15) == 0) {
        const int n8 = n >> 3;
        const uint4* m8 = reinterpret_cast<const uint4*>(m);
        auto feed8 = [&](uint4 q, int base) {
            const __half2* h = reinterpret_cast<const __half2*>(&q);
#pragma unroll
            for (int k = 0; k < 4; ++k) {
                float2 f = __half22float2(h[k]);
                feed(f.x, base + 2 * k);
                feed(f.y, base + 2 * k + 1);
            }
        };
        int i = lane;
        for (; i + 96 < n8; i += 128) {          // 4 independent 16 B loads in flight
            uint4 a = __ldg(m8 + i), c = __ldg(m8 + i + 32), d = __ldg(m8 + i + 64), e = __ldg(m8 + i + 96);
            feed8(a, 8 * i);
            feed8(c, 8 * i + 256);
            feed8(d, 8 * i + 512);
            feed8(e, 8 * i + 768);
        }
        for (; i + 32 < n8; i += 64) {
            uint4 a = __ldg(m8 + i), c = __ldg(m8 + i + 32);
            feed8(a, 8 * i);
            feed8(c, 8 * i + 256);
        }
        for (; i < n8; i += 32) feed8(__ldg(m8 + i), 8 * i);
        for (int t = (n8 << 3) + lane; t < n; t += 32) feed(__half2float(m[t]), t);
    } else {
        for (int t = lane; t < n; t += 32) feed(__half2float(m[t]), t);
    }
    if (!started) { b.v = -INFINITY; b.i = 0x7fffffff; }
    return b;
}

// python: int(a) // 2 on the float32 sum (trunc toward zero, then floor div)
__device__ __forceinline__ long long int_mid(float a, float b) {
    long long s = (long long)__fadd_rn(a, b);     // cast truncates toward zero
    long long q = s / 2;
    if ((s % 2 != 0) && (s < 0)) q -= 1;
    return q;
}

struct SegDef { signed char a, b; };
// modules/pose_estimator.py:156-166 (mirror naming); -1 chest, -2 crotch
__constant__ SegDef kSeg[11] = {{5, 6}, {-2, -1}, {5, 7}, {6, 8}, {9, 7}, {10, 8},
                                {12, 11}, {12, 14}, {11, 13}, {16, 14}, {15, 13}};

// The 11 segment lengths of one person from its 17 image-space keypoints (lane j holds joint j), called by a full warp
// (modules/pose_estimator.py:130-200).
__device__ __forceinline__ void segment_lengths(int lane, float ix, float iy, uint32_t ign_mask, double p2c, int p,
                                                float* __restrict__ lengths, double* __restrict__ torso) {
    // chest / crotch: integer midpoints (pose_estimator.py:146-153)
    const float x5 = __shfl_sync(0xffffffffu, ix, 5), y5 = __shfl_sync(0xffffffffu, iy, 5);
    const float x6 = __shfl_sync(0xffffffffu, ix, 6), y6 = __shfl_sync(0xffffffffu, iy, 6);
    const float x11 = __shfl_sync(0xffffffffu, ix, 11), y11 = __shfl_sync(0xffffffffu, iy, 11);
    const float x12 = __shfl_sync(0xffffffffu, ix, 12), y12 = __shfl_sync(0xffffffffu, iy, 12);
    const bool have_chest = !((ign_mask >> 5) & 1) && !((ign_mask >> 6) & 1);
    const bool have_crotch = !((ign_mask >> 11) & 1) && !((ign_mask >> 12) & 1);
    const SegDef sd = kSeg[lane < 11 ? lane : 0];
    const float ax = __shfl_sync(0xffffffffu, ix, sd.a < 0 ? 0 : sd.a), ay = __shfl_sync(0xffffffffu, iy, sd.a < 0 ? 0 : sd.a);
    const float bx = __shfl_sync(0xffffffffu, ix, sd.b < 0 ? 0 : sd.b), by = __shfl_sync(0xffffffffu, iy, sd.b < 0 ? 0 : sd.b);
    if (lane >= 11) return;
    float out = 0.f;
    double out_d = 0.0;
    if (lane == 1) {
        if (have_chest && have_crotch) {
            const long long dx = int_mid(x11, x12) - int_mid(x5, x6);
            const long long dy = int_mid(y11, y12) - int_mid(y5, y6);
            const double nrm = sqrt((double)dx * (double)dx + (double)dy * (double)dy);
            if (nrm > 0.0) { out_d = nrm * p2c; out = (float)out_d; }
        }
        if (torso) torso[p] = out_d;
    } else {
        const bool vis = !((ign_mask >> sd.a) & 1) && !((ign_mask >> sd.b) & 1);
        if (vis) {
            // np.linalg.norm on a float32 2-vector: sqrt(dx*dx + dy*dy), no FMA
            const float dx = __fsub_rn(ax, bx), dy = __fsub_rn(ay, by);
            const float nrm = __fsqrt_rn(__fadd_rn(__fmul_rn(dx, dx), __fmul_rn(dy, dy)));
            // value * pixel_to_cm with a weak python float -> float32 multiply
            if (nrm > 0.f) out = __fmul_rn(nrm, (float)p2c);
        }
    }
    if (lengths) lengths[(size_t)p * 11 + lane] = out;
}

// pose_estimator.py:191-200 on keypoints the caller already holds: one warp per person.
__global__ void __launch_bounds__(128)
keypoint_lengths_kernel(const float* __restrict__ kpts, const uint32_t* __restrict__ ignored,
                        const double* __restrict__ pixel_to_cm, int P, float* __restrict__ lengths,
                        double* __restrict__ torso) {
    const int p = blockIdx.x * 4 + (threadIdx.x >> 5), lane = threadIdx.x & 31;
    if (p >= P) return;
    const float x = lane < 17 ? kpts[((size_t)p * 17 + lane) * 2] : 0.f;
    const float y = lane < 17 ? kpts[((size_t)p * 17 + lane) * 2 + 1] : 0.f;
    segment_lengths(lane, x, y, ignored ? ignored[p] : 0u, pixel_to_cm[p], p, lengths, torso);
}

template <typename T>
__global__ void __launch_bounds__(kThreads)
decode_proportions_kernel(const T* __restrict__ hm, int P, int J, int Hh, int Wh,
                          const float* __restrict__ boxes, const double* __restrict__ height_cm,
                          const float* __restrict__ thr, int quarter,
                          float* __restrict__ kpts_hm, float* __restrict__ kpts_img,
                          float* __restrict__ scores, int32_t* __restrict__ idx_out,
                          uint32_t* __restrict__ ignored_out, float* __restrict__ lengths,
                          double* __restrict__ torso, const int* __restrict__ live, const double* __restrict__ Maff,
                          int crop_h, int crop_w) {
    __shared__ float s_x[kMaxJ], s_y[kMaxJ], s_v[kMaxJ];
    __shared__ int s_i[kMaxJ];
    const int p = blockIdx.x;
    if (live && p >= *live) return;                     // chained pipeline: person slots beyond the device-side count
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int n = Hh * Wh;
    for (int j = warp; j < J; j += kWarps) {
        const T* m = hm + ((size_t)p * J + j) * n;
        Best b;
        if constexpr (sizeof(T) == 4) b = scan_f32(reinterpret_cast<const float*>(m), n, lane);
        else b = scan_f16(reinterpret_cast<const __half*>(m), n, lane);
        b = warp_best(b);
        if (lane == 0) {
            // pose_estimator.py:93-98: x = idx % W, y = floor(idx / W) in float32
            // (exact below 2^24), both zeroed unless max > 0
            float x = (float)(b.i % Wh), y = (float)(b.i / Wh);
            const bool pos = b.v > 0.0f;
            if (!pos) { x = 0.f; y = 0.f; }
            if (quarter && pos) {      // public HRNet get_final_preds, not in the reference
                const int xi = b.i % Wh, yi = b.i / Wh;
                if (xi > 1 && xi < Wh - 1 && yi > 1 && yi < Hh - 1) {
                    auto at = [&](int yy, int xx) {
                        if constexpr (sizeof(T) == 4) return (float)m[yy * Wh + xx];
                        else return __half2float(m[yy * Wh + xx]);
                    };
                    const float dx = __fsub_rn(at(yi, xi + 1), at(yi, xi - 1));
                    const float dy = __fsub_rn(at(yi + 1, xi), at(yi - 1, xi));
                    x = __fadd_rn(x, 0.25f * (dx > 0.f ? 1.f : dx < 0.f ? -1.f : dx));
                    y = __fadd_rn(y, 0.25f * (dy > 0.f ? 1.f : dy < 0.f ? -1.f : dy));
                }
            }
            s_x[j] = x; s_y[j] = y; s_v[j] = b.v; s_i[j] = b.i;
        }
    }
    __syncthreads();
    if (warp != 0) return;

    const bool have_j = lane < J;
    float x = have_j ? s_x[lane] : 0.f, y = have_j ? s_y[lane] : 0.f, v = have_j ? s_v[lane] : 0.f;
    if (have_j) {
        const size_t o = (size_t)p * J + lane;
        if (kpts_hm) { kpts_hm[2 * o] = x; kpts_hm[2 * o + 1] = y; }
        if (scores) scores[o] = v;
        if (idx_out) idx_out[o] = s_i[lane];
    }
    if (!boxes) return;
    // person_det_pose_edet4_trtserver.py:151-160
    const float by1 = boxes[4 * p], bx1 = boxes[4 * p + 1], by2 = boxes[4 * p + 2], bx2 = boxes[4 * p + 3];
    const int x1 = (int)bx1, y1 = (int)by1, x2 = (int)bx2, y2 = (int)by2;   // int() truncation
    const float cw = (float)(x2 - x1), ch = (float)(y2 - y1);
    float ix = __fadd_rn(__fmul_rn(__fdiv_rn(x, (float)Wh), cw), (float)x1);
    float iy = __fadd_rn(__fmul_rn(__fdiv_rn(y, (float)Hh), ch), (float)y1);
    if (Maff) {
        // optional general inverse affine (north_star item 5): the crop's own dst->src matrix (what hbp_crop_warp sampled
        // with, rotation / aspect padding included) applied to the heatmap cell's crop coordinate
        // (u,v) = (x*crop_w/Wh, y*crop_h/Hh), in double, rounded once to float32.  The reference's box formula above stays
        // the default.
        const double* m = Maff + (size_t)p * 6;
        const double u = __ddiv_rn(__dmul_rn((double)x, (double)crop_w), (double)Wh);
        const double v = __ddiv_rn(__dmul_rn((double)y, (double)crop_h), (double)Hh);
        ix = (float)__dadd_rn(__dadd_rn(__dmul_rn(m[0], u), __dmul_rn(m[1], v)), m[2]);
        iy = (float)__dadd_rn(__dadd_rn(__dmul_rn(m[3], u), __dmul_rn(m[4], v)), m[5]);
    }
    if (have_j && kpts_img) {
        const size_t o = (size_t)p * J + lane;
        kpts_img[2 * o] = ix; kpts_img[2 * o + 1] = iy;
    }
    // :162-163  ignored iff score < T_j (false for NaN)
    const bool ign = have_j && (v < thr[lane]);
    const uint32_t ign_mask = __ballot_sync(0xffffffffu, ign);
    if (lane == 0 && ignored_out) ignored_out[p] = ign_mask;
    if (!(lengths || torso) || J != 17) return;
    // :166-168 pixel_to_cm = height_cm / (y2 - y1) in python float (double)
    segment_lengths(lane, ix, iy, ign_mask, height_cm[p] / (double)(y2 - y1), p, lengths, torso);
}

}  // namespace

int k_decode_proportions(hbp_ctx* ctx, const void* hm, int dtype, int P, int J, int Hh, int Wh,
                         const float* boxes, const double* height_cm, const float* thr, int quarter,
                         float* kpts_hm, float* kpts_img, float* scores, int32_t* idx,
                         uint32_t* ignored, float* lengths, double* torso, const int* live, const double* Maff,
                         int crop_h, int crop_w) {
    if (P <= 0) return HBP_OK;
    if (dtype == HBP_F32)
        decode_proportions_kernel<float><<<P, kThreads, 0, ctx->stream>>>(
            (const float*)hm, P, J, Hh, Wh, boxes, height_cm, thr, quarter, kpts_hm, kpts_img, scores,
            idx, ignored, lengths, torso, live, Maff, crop_h, crop_w);
    else
        decode_proportions_kernel<__half><<<P, kThreads, 0, ctx->stream>>>(
            (const __half*)hm, P, J, Hh, Wh, boxes, height_cm, thr, quarter, kpts_hm, kpts_img, scores,
            idx, ignored, lengths, torso, live, Maff, crop_h, crop_w);
    HBP_LAUNCH_CHECK(ctx);
    return HBP_OK;
}

int k_keypoint_lengths(hbp_ctx* ctx, const float* kpts, const uint32_t* ignored, const double* pixel_to_cm, int P,
                       float* lengths, double* torso) {
    if (P <= 0) return HBP_OK;
    keypoint_lengths_kernel<<<(P + 3) / 4, 128, 0, ctx->stream>>>(kpts, ignored, pixel_to_cm, P, lengths, torso);
    HBP_LAUNCH_CHECK(ctx);
    return HBP_OK;
}
