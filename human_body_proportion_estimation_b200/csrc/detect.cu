// K2 / K3 / K3b: detector-head post-processing.
//
// Reference (relative to the reference's human_body_length_est/ unless noted):
//   raw-head decode ....... obj_det_yolov5_onnx.py:123-169
//   official NMS .......... modules/onnx_utils.py:125-222 -> torchvision.ops.nms (:205)
//   legacy NMS ............ modules/onnx_utils.py:8-95
//   scale/clip coords ..... modules/onnx_utils.py:238-266
//   EfficientDet filter ... models/conv.py:22-57 (reference root)
//
// NMS pipeline, all on the device, no host round trip:
//   1. filter    obj > conf per row, then for the rows that pass: cls*obj, first-max class (torch.max order), conf > thr,
//                class filter, append to the candidate list (order arbitrary: step 2 sorts).  Two forms with the same
//                per-row arithmetic: up to 8 frames every CTA streams its 256 rows through shared memory with coalesced
//                16-byte loads (yolo_filter_stream_kernel: one memory round trip for the whole head); large batches read
//                only the objectness sector of every row and the flagged rows (yolo_filter_kernel: a tenth of the bytes).
//   2. rank      stable order = (conf desc, source row asc); one warp per candidate counts the candidates that precede
//                it (its lanes split the others) and scatters it to its rank.
//   3. mask      bit (i,j), j>i, set iff IoU(i,j) > thr on the class-offset
//                boxes; one warp per (row, 32-column word) builds the word with
//                a ballot.  IoU in torchvision's operation order with explicit
//                round-to-nearest intrinsics (no FMA contraction):
//                inter/(area_i + area_j - inter), compared as double.
//   4. sweep + gather, one launch, one CTA per image: the image's mask (n x ceil(n/32) words, rows stored with the
//                image's own stride) is staged in shared memory with one linear copy when it fits (n <= ~1260), one warp
//                walks the words -- lane ww holds the removed bits of word ww in a register, every KEPT box costs one
//                shared-memory read per lane and a shuffle -- then all threads write [x1,y1,x2,y2,conf,cls] of the kept
//                boxes (un-offset).  Stops after max_det keeps.
// Grids are sized on the host without knowing the candidate count (the kernels read it on the device and stride); the
// launches are chained with programmatic dependent launch.  One 25200 x 85 head, 64 persons kept: 131 us (round 1) -> 38 us
// (profiles/r02_nms_kernels.md).
// The legacy variant reuses 2-4 with key (class asc, obj desc, row asc), the
// +1-pixel IoU, "suppress unless iou < thr" and same-class-only suppression.
#include "hbp_internal.cuh"
#include <algorithm>

namespace {

// programmatic dependent launch (the NMS kernels form a chain of short launches: the next one's CTAs are scheduled while
// the previous one drains, and wait here for its results)
__device__ __forceinline__ void pdl_trigger() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }
__device__ __forceinline__ void pdl_sync() { asm volatile("griddepcontrol.wait;" ::: "memory"); }

template <typename... KArgs, typename... Args>
void launch_chained(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t st, Args... args) {
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = grid; cfg.blockDim = block; cfg.dynamicSmemBytes = smem; cfg.stream = st;
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    at[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = at; cfg.numAttrs = 1;
    cudaLaunchKernelEx(&cfg, kernel, KArgs(args)...);
}

__constant__ float kAnchors[3][6] = {{116, 90, 156, 198, 373, 326},
                                     {30, 61, 62, 45, 59, 119},
                                     {10, 13, 16, 30, 33, 23}};

__device__ __forceinline__ float sigmoidf_(float x) { return __fdiv_rn(1.f, __fadd_rn(1.f, expf(-x))); }

// ---- raw head decode ---------------------------------------------------------
// in: (B,3,S,S,E) ; out rows at row_off + ((a*S + gy)*S + gx) of an image with
// total_rows rows.  One thread per element, coalesced on E.
__global__ void __launch_bounds__(256)
yolo_decode_kernel(const float* __restrict__ in, float* __restrict__ out, int B, int S, int E,
                   int level, float stride_w, float stride_h, int row_off, int total_rows) {
    const size_t per_img = (size_t)3 * S * S * E;
    const size_t total = per_img * B;
    const size_t step = (size_t)gridDim.x * blockDim.x;
    for (size_t t = (size_t)blockIdx.x * blockDim.x + threadIdx.x; t < total; t += step) {
        const int b = (int)(t / per_img);
        const size_t r = t % per_img;
        const int e = (int)(r % E);
        const size_t cell = r / E;              // (a*S + i)*S + j
        const int j = (int)(cell % S), i = (int)((cell / S) % S), a = (int)(cell / ((size_t)S * S));
        const float s = sigmoidf_(in[t]);
        float v = s;
        // the reference builds grid_x/grid_y with meshgrid(arange(shape[2]), arange(shape[3]))
        // and adds them to a (...,S,S) tensor: grid_x varies along the LAST axis (j).
        if (e == 0) v = __fmul_rn(__fadd_rn(__fsub_rn(__fmul_rn(s, 2.f), 0.5f), (float)j), stride_w);
        else if (e == 1) v = __fmul_rn(__fadd_rn(__fsub_rn(__fmul_rn(s, 2.f), 0.5f), (float)i), stride_h);
        else if (e == 2 || e == 3) {
            const float q = __fmul_rn(s, 2.f);
            v = __fmul_rn(__fmul_rn(q, q), kAnchors[level][2 * a + (e - 2)]);
        }
        out[((size_t)b * total_rows + row_off + cell) * E + e] = v;
    }
}

// ---- candidates ---------------------------------------------------------------
constexpr int kFilterUnroll = 4;

struct Cand {           // 32 bytes
    float x1, y1, x2, y2;
    float conf;         // official: obj*cls ; legacy: obj
    float cls;
    float aux;          // legacy: class confidence
    int src;            // source row (tie-break)
};

__global__ void __launch_bounds__(256)
yolo_filter_kernel(const float* __restrict__ pred, int N, int nc, float conf_thres, int legacy,
                   const int* __restrict__ classes, int n_classes, Cand* __restrict__ cand,
                   int* __restrict__ cand_count, int cap) {
    pdl_trigger();
    const int b = blockIdx.y;
    const int E = 5 + nc;
    const int lane = threadIdx.x & 31;
    const int warp_global = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int row00 = warp_global * (32 * kFilterUnroll);
    if (row00 >= N) return;
    const float* __restrict__ base = pred + (size_t)b * N * E;
    // kFilterUnroll groups of 32 rows per warp: all objectness loads (one 32-byte sector of a 340-byte row each) are in
    // flight before the first ballot -- the kernel is bound by the latency of these strided loads, not by bytes
    float objs[kFilterUnroll];
#pragma unroll
    for (int u = 0; u < kFilterUnroll; ++u) {
        const int my_row = row00 + 32 * u + lane;
        objs[u] = my_row < N ? __ldg(base + (size_t)my_row * E + 4) : 0.f;
    }
    // One LANE per flagged row (obj > conf).  A row is 5 + nc floats = up to 12 sectors: the lane first touches every sector
    // of all its flagged rows (independent loads, one L2 round trip for all of them), then walks the class scores out of
    // L1 -- walking them cold cost one L2 round trip per sector and row (22 us for a 25200-row head: ncu, r02af).
    bool flags[kFilterUnroll];
    float keep_alive = 0.f;
#pragma unroll
    for (int u = 0; u < kFilterUnroll; ++u) {
        const int row = row00 + 32 * u + lane;
        flags[u] = (row < N) && (legacy ? (objs[u] >= conf_thres) : (objs[u] > conf_thres));
        if (flags[u]) {
            const float* __restrict__ r = base + (size_t)row * E;
            for (int j = 8; j < E; j += 8) keep_alive += __ldg(r + j);
            keep_alive += __ldg(r + E - 1);
        }
    }
    float bests[kFilterUnroll];
    int best_js[kFilterUnroll];
    unsigned okm[kFilterUnroll];
    int total = 0;
#pragma unroll
    for (int u = 0; u < kFilterUnroll; ++u) {
        const int row = row00 + 32 * u + lane;
        const float* __restrict__ r = base + (size_t)row * E;
        const float o = objs[u];
        float best = -INFINITY;
        int best_j = 0x7fffffff;
        bool ok = false;
        if (flags[u]) {
            // first-max over classes of (cls*obj) [official] or cls [legacy]; torch.max: NaN propagates as the max, first index on ties
#pragma unroll 8
            for (int j = 0; j < nc; ++j) {
                float c = __ldg(r + 5 + j);
                if (!legacy) c = __fmul_rn(c, o);
                if (j == 0) { best = c; best_j = 0; }
                else if (c > best || (c != c && best == best)) { best = c; best_j = j; }
            }
            if (keep_alive == 1.2345e38f) best = keep_alive;          // (never true: keeps the sector touches above alive)
            ok = legacy ? true : (best > conf_thres);
            if (ok && n_classes > 0) {
                ok = false;
                for (int k = 0; k < n_classes; ++k) ok |= ((float)best_j == (float)classes[k]);
            }
        }
        bests[u] = best; best_js[u] = best_j;
        okm[u] = __ballot_sync(0xffffffffu, ok);
        total += __popc(okm[u]);
    }
    // ONE atomic per warp (128 rows): ~800 per-group atomics on the image's single counter serialise at its L2 slice
    // (~27 cycles each) and were most of the kernel's 22 us
    if (total == 0) return;
    int slot0 = 0;
    if (lane == 0) slot0 = atomicAdd(cand_count + b, total);
    slot0 = __shfl_sync(0xffffffffu, slot0, 0);
#pragma unroll
    for (int u = 0; u < kFilterUnroll; ++u) {
        const int slot = slot0 + __popc(okm[u] & ((1u << lane) - 1u));
        slot0 += __popc(okm[u]);
        if (((okm[u] >> lane) & 1u) && slot < cap) {
            const int row = row00 + 32 * u + lane;
            const float* __restrict__ r = base + (size_t)row * E;
            const float cx = __ldg(r), cy = __ldg(r + 1), w = __ldg(r + 2), h = __ldg(r + 3);
            const float hw = __fdiv_rn(w, 2.f), hh = __fdiv_rn(h, 2.f);
            Cand c;
            c.x1 = __fsub_rn(cx, hw); c.y1 = __fsub_rn(cy, hh);
            c.x2 = __fadd_rn(cx, hw); c.y2 = __fadd_rn(cy, hh);
            c.conf = legacy ? objs[u] : bests[u];
            c.cls = (float)best_js[u];
            c.aux = bests[u];
            c.src = row;
            cand[(size_t)b * cap + slot] = c;
        }
    }
}

// does candidate a come before candidate b in the output order?
__device__ __forceinline__ bool precedes(float ac, float acls, int as, float bc, float bcls, int bs, int legacy) {
    if (legacy && acls != bcls) return acls < bcls;
    if (ac != bc) return ac > bc;
    return as < bs;
}

struct SortedBox {      // 32 bytes
    float x1, y1, x2, y2;   // boxes as suppressed on (class-offset for the official path)
    float conf, cls, aux;
    int cand;               // index into the candidate list
};

// One WARP per candidate: its lanes split the other candidates, count the ones that precede it and reduce -- n / 8 CTAs
// of eight warps (150 for the 1200 candidates of a 25200-row head) instead of n / 256 CTAs whose every thread walked all n.
constexpr int kRankWarps = 8;
__global__ void __launch_bounds__(32 * kRankWarps)
rank_scatter_kernel(const Cand* __restrict__ cand, const int* __restrict__ cand_count, int cap,
                    int legacy, int max_nms, float max_wh, SortedBox* __restrict__ sorted,
                    int* __restrict__ n_sorted, int* __restrict__ status) {
    pdl_trigger();
    pdl_sync();
    const int b = blockIdx.y;
    const int n_all = cand_count[b];
    const int n = min(n_all, cap);
    const Cand* __restrict__ c = cand + (size_t)b * cap;
    const int lane = threadIdx.x & 31;
    if (blockIdx.x == 0 && threadIdx.x == 0) {
        n_sorted[b] = min(n, max_nms);
        if (status && n_all > cap) atomicOr(status, 1);          // candidates beyond the capacity were dropped
    }
    // (the grid is sized on the host without knowing n: the warps stride over the candidates)
    for (int i = blockIdx.x * kRankWarps + (threadIdx.x >> 5); i < n; i += gridDim.x * kRankWarps) {
    const Cand me = c[i];
    int rank = 0;
    for (int j = lane; j < n; j += 32) {
        const float4 q = __ldg(reinterpret_cast<const float4*>(c + j) + 1);      // conf, cls, aux, src
        rank += precedes(q.x, q.y, __float_as_int(q.w), me.conf, me.cls, me.src, legacy) ? 1 : 0;
    }
    rank = __reduce_add_sync(0xffffffffu, rank);
    if (lane == 0 && rank < max_nms) {
        SortedBox s;
        // onnx_utils.py:202-204: boxes + cls * max_wh (float32 mul, float32 add)
        const float off = legacy ? 0.f : __fmul_rn(me.cls, max_wh);
        s.x1 = __fadd_rn(me.x1, off); s.y1 = __fadd_rn(me.y1, off);
        s.x2 = __fadd_rn(me.x2, off); s.y2 = __fadd_rn(me.y2, off);
        s.conf = me.conf; s.cls = me.cls; s.aux = me.aux; s.cand = i;
        sorted[(size_t)b * cap + rank] = s;
    }
    }
}

__device__ __forceinline__ bool suppresses(const SortedBox& a, const SortedBox& b, double thr, int legacy) {
    if (!legacy) {
        // torchvision nms_kernel_impl<float>
        const float ia = __fmul_rn(__fsub_rn(a.x2, a.x1), __fsub_rn(a.y2, a.y1));
        const float ja = __fmul_rn(__fsub_rn(b.x2, b.x1), __fsub_rn(b.y2, b.y1));
        const float w = fmaxf(0.f, __fsub_rn(fminf(a.x2, b.x2), fmaxf(a.x1, b.x1)));
        const float h = fmaxf(0.f, __fsub_rn(fminf(a.y2, b.y2), fmaxf(a.y1, b.y1)));
        const float inter = __fmul_rn(w, h);
        const float ovr = __fdiv_rn(inter, __fsub_rn(__fadd_rn(ia, ja), inter));
        return (double)ovr > thr;
    }
    if (a.cls != b.cls) return false;
    // onnx_utils.py:23-34: +1 pixel convention, +1e-16; keep while iou < thr
    const float ix1 = fmaxf(a.x1, b.x1), iy1 = fmaxf(a.y1, b.y1);
    const float ix2 = fminf(a.x2, b.x2), iy2 = fminf(a.y2, b.y2);
    const float inter = __fmul_rn(fmaxf(__fadd_rn(__fsub_rn(ix2, ix1), 1.f), 0.f),
                                  fmaxf(__fadd_rn(__fsub_rn(iy2, iy1), 1.f), 0.f));
    const float aa = __fmul_rn(__fadd_rn(__fsub_rn(a.x2, a.x1), 1.f), __fadd_rn(__fsub_rn(a.y2, a.y1), 1.f));
    const float ab = __fmul_rn(__fadd_rn(__fsub_rn(b.x2, b.x1), 1.f), __fadd_rn(__fsub_rn(b.y2, b.y1), 1.f));
    const float iou = __fdiv_rn(inter, __fadd_rn(__fsub_rn(__fadd_rn(aa, ab), inter), 1e-16f));
    return !(iou < (float)thr);
}

// mask[(i*words + w)] bit l  <=>  box i suppresses box 32*w+l  (only j > i)
__global__ void __launch_bounds__(256)
nms_mask_kernel(const SortedBox* __restrict__ sorted, const int* __restrict__ n_sorted, int cap,
                int words_cap, size_t mask_img_stride, double thr, int legacy, uint32_t* __restrict__ mask) {
    pdl_trigger();
    pdl_sync();
    const int b = blockIdx.z;
    const int n = n_sorted[b];
    const int words = (n + 31) >> 5;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const SortedBox* __restrict__ s = sorted + (size_t)b * cap;
    for (int w = blockIdx.x; w < words; w += gridDim.x) {      // column word (grid sized without knowing n)
    const int j = w * 32 + lane;
    SortedBox bj{};
    if (j < n) bj = s[j];
    uint32_t* __restrict__ m = mask + (size_t)b * mask_img_stride;
    // rows are stored with the image's own stride (`words`, known on the device only), so that the sweep kernel can stage
    // the whole mask with one linear copy; words before the diagonal are written as zero
    for (int i = blockIdx.y * 8 + warp; i < n; i += gridDim.y * 8) {
        if (w * 32 + 31 < i) {
            if (lane == 0) m[(size_t)i * words + w] = 0u;
            continue;
        }
        const SortedBox bi = s[i];
        const bool bit = (j < n) && (j > i) && suppresses(bi, bj, thr, legacy);
        const uint32_t word = __ballot_sync(0xffffffffu, bit);
        if (lane == 0) m[(size_t)i * words + w] = word;
    }
    }
}

constexpr int kMaxWords = 1024;     // 32768 boxes

__global__ void scale_coords_kernel(float* __restrict__ boxes, int n, float pad_x, float pad_y, float gain,
                                    float w0, float h0) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    float* b = boxes + 4 * (size_t)i;
    // onnx_utils.py:262-265 then clip_coords :238-249
    const float x1 = __fdiv_rn(__fsub_rn(b[0], pad_x), gain), y1 = __fdiv_rn(__fsub_rn(b[1], pad_y), gain);
    const float x2 = __fdiv_rn(__fsub_rn(b[2], pad_x), gain), y2 = __fdiv_rn(__fsub_rn(b[3], pad_y), gain);
    b[0] = fminf(fmaxf(x1, 0.f), w0); b[1] = fminf(fmaxf(y1, 0.f), h0);
    b[2] = fminf(fmaxf(x2, 0.f), w0); b[3] = fminf(fmaxf(y2, 0.f), h0);
}

// models/conv.py:22-57: one warp per frame, ordered compaction with ballots.
__global__ void edet_filter_kernel(const float* __restrict__ boxes, const float* __restrict__ scores,
                                   const float* __restrict__ classes, int K, float person_class, float thr,
                                   float xe, float ye, float hf, float wf, int max_persons,
                                   float* __restrict__ out_boxes, int* __restrict__ out_count) {
    const int f = blockIdx.x, lane = threadIdx.x;
    int kept = 0;
    for (int k0 = 0; k0 < K && kept < max_persons; k0 += 32) {
        const int k = k0 + lane;
        bool ok = false;
        if (k < K) ok = (classes[(size_t)f * K + k] == person_class) && (scores[(size_t)f * K + k] >= thr);
        const unsigned bal = __ballot_sync(0xffffffffu, ok);
        const int pos = kept + __popc(bal & ((1u << lane) - 1u));
        if (ok && pos < max_persons) {
            const float* b = boxes + ((size_t)f * K + k) * 4;
            const float y1 = fminf(fmaxf(__fsub_rn(b[0], ye), 0.f), hf), x1 = fminf(fmaxf(__fsub_rn(b[1], xe), 0.f), wf);
            const float y2 = fminf(fmaxf(__fadd_rn(b[2], ye), 0.f), hf), x2 = fminf(fmaxf(__fadd_rn(b[3], xe), 0.f), wf);
            float* o = out_boxes + ((size_t)f * max_persons + pos) * 4;
            o[0] = __fdiv_rn(y1, hf); o[1] = __fdiv_rn(x1, wf); o[2] = __fdiv_rn(y2, hf); o[3] = __fdiv_rn(x2, wf);
        }
        kept += __popc(bal);
    }
    if (lane == 0) out_count[f] = min(kept, max_persons);
}

// Sweep + gather in ONE launch, one CTA per image.  When the n x ceil(n/32) words of the image's mask fit the CTA's shared
// memory (n <= ~1260) they are staged there first with coalesced loads: the sweep over the words is a chain of dependent
// reads (diagonal word -> resolve 32 boxes -> OR the kept rows -> next word), which cost a global-memory round trip per
// word and per kept row from L2 (28 words x ~1 us for the 880 candidates of a 25200-row head).  The 32 diagonal words of a
// word go through shared memory, so the box-by-box resolution is register arithmetic on broadcast reads instead of 32
// dependent shuffles.  Larger n run the same code on the mask in global memory.
constexpr int kSgThreads = 512;
constexpr uint32_t kSgSmemWords = 50 * 1024;        // 200 KB of staged mask
__global__ void __launch_bounds__(kSgThreads)
nms_sweep_gather_kernel(const uint32_t* __restrict__ mask, const int* __restrict__ n_sorted, int words_cap,
                        size_t mask_img_stride, int max_keep, int* __restrict__ keep, int* __restrict__ keep_count,
                        const SortedBox* __restrict__ sorted, const Cand* __restrict__ cand, int cap,
                        const int* __restrict__ cand_count, int legacy, float* __restrict__ out, int* __restrict__ out_count, int dbg) {
    extern __shared__ uint32_t s_m[];
    const long long t_start = dbg ? clock64() : 0;
    long long t_staged = 0, t_swept = 0;
    __shared__ uint32_t s_removed[kMaxWords];
    __shared__ uint32_t s_diag[32];
    __shared__ int s_kept;
    pdl_sync();
    const int b = blockIdx.x, tid = threadIdx.x;
    const int n = n_sorted[b];
    const int words = (n + 31) >> 5;
    const uint32_t* __restrict__ m = mask + (size_t)b * mask_img_stride;
    const bool staged = (size_t)n * (size_t)words <= (size_t)kSgSmemWords;
    if (staged) {
        // the image's mask is n x words contiguous words (nms_mask_kernel): one linear copy, 16 bytes per thread and step
        const int total = n * words;
        const int t4 = ((reinterpret_cast<uintptr_t>(m) & 15) == 0) ? (total >> 2) : 0;
        const uint4* __restrict__ m4 = reinterpret_cast<const uint4*>(m);
        uint4* s4 = reinterpret_cast<uint4*>(s_m);
        for (int k = tid; k < t4; k += kSgThreads) s4[k] = __ldg(m4 + k);
        for (int k = (t4 << 2) + tid; k < total; k += kSgThreads) s_m[k] = __ldg(m + k);
    }
    for (int w = tid; w < words; w += kSgThreads) s_removed[w] = 0;
    if (tid == 0) s_kept = 0;
    __syncthreads();
    if (dbg) t_staged = clock64();
    // The sweep itself is one warp: every step depends on the one before it, block barriers would only add latency.
    // Per word: the 32 diagonal words (bits j > i inside the word) go through shared memory; the next kept box is always the
    // lowest bit that is neither removed nor decided yet, so the loop runs once per KEPT box (2-3 per word for a detector
    // head), not once per box; then the lanes OR the kept rows into the removed words behind this one.
    if (tid < 32 && staged && words <= 32) {
        // up to 1024 boxes: lane ww keeps the removed bits of word ww in a register.  Keeping box (w, l) costs one
        // conflict-free shared-memory read per lane (row's word ww) and one shuffle (the current word's new removed bits):
        // ~90 cycles per kept box instead of three dependent passes (~350)
        const int lane = tid;
        uint32_t myrem = 0;
        int kept = 0;
        for (int w = 0; w < words && kept < max_keep; ++w) {
            uint32_t decided = __shfl_sync(0xffffffffu, myrem, w);
            if (n - w * 32 < 32) decided |= ~0u << (n - w * 32);            // padding bits
            while (~decided != 0u && kept < max_keep) {
                const int l = __ffs(~decided) - 1;
                const int row = w * 32 + l;
                if (lane == 0) keep[(size_t)b * max_keep + kept] = row;
                ++kept;
                if (lane < words) myrem |= s_m[row * words + lane];     // (words before the diagonal are stored as zero)
                decided |= __shfl_sync(0xffffffffu, myrem, w) | (1u << l);
            }
        }
        if (lane == 0) s_kept = kept;
    } else if (tid < 32) {
        const int lane = tid;
        int kept = 0;
        for (int w = 0; w < words && kept < max_keep; ++w) {
            const int i = w * 32 + lane;
            s_diag[lane] = (i < n) ? (staged ? s_m[i * words + w] : m[(size_t)i * words + w]) : 0u;
            __syncwarp();
            uint32_t removed = s_removed[w];
            if (n - w * 32 < 32) removed |= ~0u << (n - w * 32);            // padding bits
            uint32_t keepbits = 0;
            while (~removed != 0u && kept < max_keep) {
                const int l = __ffs(~removed) - 1;
                keepbits |= 1u << l;
                ++kept;
                removed |= s_diag[l] | (1u << l);
            }
            // (keepbits / kept are warp-uniform: every lane ran the same loop on broadcast values)
            if (keepbits) {
                if (lane == 0) {
                    int kk = kept - __popc(keepbits);
                    for (uint32_t kb = keepbits; kb; kb &= kb - 1) keep[(size_t)b * max_keep + kk++] = w * 32 + __ffs(kb) - 1;
                }
                for (int ww = w + 1 + lane; ww < words; ww += 32) {
                    uint32_t acc = 0;
                    for (uint32_t t = keepbits; t; t &= t - 1) {
                        const int row = w * 32 + __ffs(t) - 1;
                        acc |= staged ? s_m[row * words + ww] : m[(size_t)row * words + ww];
                    }
                    s_removed[ww] |= acc;
                }
            }
            __syncwarp();
        }
        if (lane == 0) s_kept = kept;
    }
    __syncthreads();
    if (dbg) t_swept = clock64();
    // gather (keep[] was written by thread 0 before block barriers)
    const int nk = s_kept;
    const int width = legacy ? 7 : 6;
    for (int k = tid; k < nk; k += kSgThreads) {
        const SortedBox sb = sorted[(size_t)b * cap + keep[(size_t)b * max_keep + k]];
        const Cand c = cand[(size_t)b * cap + sb.cand];
        float* o = out + ((size_t)b * max_keep + k) * width;
        o[0] = c.x1; o[1] = c.y1; o[2] = c.x2; o[3] = c.y2;
        if (legacy) { o[4] = c.conf; o[5] = c.aux; o[6] = c.cls; }
        else { o[4] = c.conf; o[5] = c.cls; }
    }
    if (dbg && tid == 0) printf("[nms sweep] n=%d words=%d staged=%d kept=%d | cycles: staging %lld, sweep %lld, gather %lld\n", n, words, (int)staged, nk, t_staged - t_start, t_swept - t_staged, clock64() - t_swept);
    if (tid == 0) {
        keep_count[b] = nk;
        out_count[b] = (legacy && cand_count[b] == 0) ? -1 : nk;
    }
}

// Latency form of the candidate filter (a few frames): a CTA copies its kStreamRows rows -- one contiguous span of the head --
// into shared memory with coalesced 16-byte loads (all of them in flight at once: one memory round trip for the whole head
// instead of a chain of strided sector reads per flagged row), then one thread per row tests the objectness and, for the ~4 %
// of rows that pass, walks the class scores out of shared memory (row stride 5 + nc words: conflict-free for odd strides).
// One counter atomic per CTA.  It moves the whole head (8.6 MB for 25200 x 85) where the sector-wise kernel above moves a
// tenth of it, which is why large batches keep the other one.
constexpr int kStreamRows = 256;
__global__ void __launch_bounds__(kStreamRows)
yolo_filter_stream_kernel(const float* __restrict__ pred, int N, int nc, float conf_thres, int legacy,
                          const int* __restrict__ classes, int n_classes, Cand* __restrict__ cand,
                          int* __restrict__ cand_count, int cap) {
    extern __shared__ __align__(16) unsigned char s_rows_raw[];
    __shared__ int s_warp_n[kStreamRows / 32];
    __shared__ int s_base;
    pdl_trigger();
    const int b = blockIdx.y, tid = threadIdx.x, lane = tid & 31, wrp = tid >> 5;
    const int E = 5 + nc;
    const int r0 = blockIdx.x * kStreamRows;
    const int rows = min(kStreamRows, N - r0);
    const unsigned char* gbytes = reinterpret_cast<const unsigned char*>(pred + ((size_t)b * N + r0) * E);
    const size_t span = (size_t)rows * E * sizeof(float);
    const size_t lead = reinterpret_cast<uintptr_t>(gbytes) & 15;              // the span keeps its 16-byte phase in shared memory
    const unsigned char* g0 = gbytes - lead;
    const size_t vecs = (lead + span + 15) / 16;
    // (the aligned vectors may reach up to 15 bytes before / after the span: still inside the allocation unless the span is
    // the very first / last bytes of it -- those two vectors are read bytewise)
    const unsigned char* t_begin = reinterpret_cast<const unsigned char*>(pred);
    const unsigned char* t_end = reinterpret_cast<const unsigned char*>(pred + (size_t)gridDim.y * N * E);
    for (size_t v = tid; v < vecs; v += kStreamRows) {
        const unsigned char* ga = g0 + 16 * v;
        uint4 q = make_uint4(0, 0, 0, 0);
        if (ga >= t_begin && ga + 16 <= t_end) q = __ldg(reinterpret_cast<const uint4*>(ga));
        else {
            unsigned char* qb = reinterpret_cast<unsigned char*>(&q);
            for (int k = 0; k < 16; ++k) if (ga + k >= t_begin && ga + k < t_end) qb[k] = __ldg(ga + k);
        }
        *reinterpret_cast<uint4*>(s_rows_raw + 16 * v) = q;
    }
    __syncthreads();
    const float* srow = reinterpret_cast<const float*>(s_rows_raw + lead) + (size_t)tid * E;
    bool ok = false;
    float best = -INFINITY, o = 0.f;
    int best_j = 0x7fffffff;
    if (tid < rows) {
        o = srow[4];
        if (legacy ? (o >= conf_thres) : (o > conf_thres)) {
            // first-max over classes of (cls*obj) [official] or cls [legacy]; torch.max: NaN propagates as the max, first index on ties
            for (int j = 0; j < nc; ++j) {
                float c = srow[5 + j];
                if (!legacy) c = __fmul_rn(c, o);
                if (j == 0) { best = c; best_j = 0; }
                else if (c > best || (c != c && best == best)) { best = c; best_j = j; }
            }
            ok = legacy ? true : (best > conf_thres);
            if (ok && n_classes > 0) {
                ok = false;
                for (int k = 0; k < n_classes; ++k) ok |= ((float)best_j == (float)classes[k]);
            }
        }
    }
    const unsigned okm = __ballot_sync(0xffffffffu, ok);
    if (lane == 0) s_warp_n[wrp] = __popc(okm);
    __syncthreads();
    if (tid == 0) {
        int tot = 0;
        for (int w = 0; w < kStreamRows / 32; ++w) { const int c = s_warp_n[w]; s_warp_n[w] = tot; tot += c; }
        s_base = tot ? atomicAdd(cand_count + b, tot) : 0;
    }
    __syncthreads();
    if (ok) {
        const int slot = s_base + s_warp_n[wrp] + __popc(okm & ((1u << lane) - 1u));
        if (slot < cap) {
            const float cx = srow[0], cy = srow[1], w = srow[2], h = srow[3];
            const float hw = __fdiv_rn(w, 2.f), hh = __fdiv_rn(h, 2.f);
            Cand c;
            c.x1 = __fsub_rn(cx, hw); c.y1 = __fsub_rn(cy, hh);
            c.x2 = __fadd_rn(cx, hw); c.y2 = __fadd_rn(cy, hh);
            c.conf = legacy ? o : best;
            c.cls = (float)best_j;
            c.aux = best;
            c.src = r0 + tid;
            cand[(size_t)b * cap + slot] = c;
        }
    }
}

void launch_filter(hbp_ctx* ctx, const float* pred, int B, int N, int nc, float conf, int legacy, const int* classes,
                   int n_classes, Cand* cand, int* cand_count, int cap) {
    // a few frames: two warps per CTA (~100 CTAs for one 25200-row head: the kernel is a chain of strided-load round
    // trips, it needs SMs, not threads); large batches: eight
    static const int stream_max_b = getenv("HBP_FILTER_STREAM_MAXB") ? atoi(getenv("HBP_FILTER_STREAM_MAXB")) : 8;
    const size_t smem = (size_t)kStreamRows * (5 + nc) * sizeof(float) + 32;
    bool stream_ok = B <= stream_max_b && smem <= 200 * 1024;
    if (stream_ok && !(ctx->attr_flags & ATTR_FILTER)) {
        if (cudaFuncSetAttribute(yolo_filter_stream_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024) == cudaSuccess) ctx->attr_flags |= ATTR_FILTER;
        else { cudaGetLastError(); stream_ok = false; }
    }
    if (stream_ok) {
        dim3 grid((N + kStreamRows - 1) / kStreamRows, B);
        yolo_filter_stream_kernel<<<grid, kStreamRows, smem, ctx->stream>>>(pred, N, nc, conf, legacy, classes, n_classes, cand, cand_count, cap);
        return;
    }
    const int warps = (N + 32 * kFilterUnroll - 1) / (32 * kFilterUnroll);
    const int wpb = B >= 16 ? 8 : 2;
    dim3 grid((warps + wpb - 1) / wpb, B);
    yolo_filter_kernel<<<grid, 32 * wpb, 0, ctx->stream>>>(pred, N, nc, conf, legacy, classes, n_classes, cand, cand_count, cap);
}

// rank + scatter, suppression mask, sweep + gather for up to `n_max` sorted candidates per image (the kernels read the
// actual counts on the device: nothing here depends on them)
int launch_nms_tail(hbp_ctx* ctx, const Cand* cand, const int* cand_count, SortedBox* sorted, int* n_sorted, uint32_t* mask,
                    int* keep, int* keep_count, int B, int cap, int n_max, int words_cap, size_t mask_img_stride, double thr,
                    int legacy, int max_nms, int max_keep, int* status, float* out_det, int* out_count) {
    if (!(ctx->attr_flags & ATTR_NMS)) {            // (function attributes are per device: one flag per context, not a process-wide static)
        HBP_CUDA(cudaFuncSetAttribute(nms_sweep_gather_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(kSgSmemWords * 4)));
        ctx->attr_flags |= ATTR_NMS;
    }
    const int rank_blocks = std::min((n_max + kRankWarps - 1) / kRankWarps, 2 * ctx->sm_count);
    launch_chained(rank_scatter_kernel, dim3(std::max(rank_blocks, 1), B), dim3(32 * kRankWarps), 0, ctx->stream, cand, cand_count, cap, legacy, max_nms,
                   4096.f, sorted, n_sorted, status);
    HBP_LAUNCH_CHECK(ctx);
    const int words = (n_max + 31) / 32;
    dim3 g2(std::max(std::min(words, 64), 1), 16, B);
    launch_chained(nms_mask_kernel, g2, dim3(256), 0, ctx->stream, (const SortedBox*)sorted, (const int*)n_sorted, cap, words_cap, mask_img_stride, thr, legacy, mask);
    HBP_LAUNCH_CHECK(ctx);
    const size_t need = (size_t)n_max * (size_t)words * 4;
    const size_t dyn = std::min(need, (size_t)kSgSmemWords * 4);
    launch_chained(nms_sweep_gather_kernel, dim3(B), dim3(kSgThreads), dyn, ctx->stream, (const uint32_t*)mask, (const int*)n_sorted, words_cap, mask_img_stride,
                   max_keep, keep, keep_count, (const SortedBox*)sorted, cand, cap, cand_count, legacy, out_det, out_count, getenv("HBP_NMS_DBG") ? 1 : 0);
    HBP_LAUNCH_CHECK(ctx);
    return HBP_OK;
}

int run_nms(hbp_ctx* ctx, const float* pred, int B, int N, int nc, float conf, double thr,
            const int* classes, int n_classes, int max_keep, int legacy, float* out_det, int* out_count) {
    const int cap = N;                                   // every row can be a candidate
    const int max_nms = legacy ? N : 30000;              // onnx_utils.py:143
    if (N > kMaxWords * 32) { hbp_set_error("NMS supports at most %d rows per image", kMaxWords * 32); return HBP_ERR_INVALID; }
    Cand* cand = (Cand*)hbp_scratch(ctx, SC_NMS_CAND, (size_t)B * cap * sizeof(Cand));
    SortedBox* sorted = (SortedBox*)hbp_scratch(ctx, SC_NMS_SORTED, (size_t)B * cap * sizeof(SortedBox));
    // misc: cand_count[B] | n_sorted[B] | keep_count[B] | keep[B*max_keep]
    int* misc = (int*)hbp_scratch(ctx, SC_NMS_MISC, ((size_t)3 * B + (size_t)B * max_keep) * sizeof(int));
    if (!cand || !sorted || !misc) return HBP_ERR_NOMEM;
    int* cand_count = misc, *n_sorted = misc + B, *keep_count = misc + 2 * B, *keep = misc + 3 * B;
    HBP_CUDA(cudaMemsetAsync(misc, 0, (size_t)3 * B * sizeof(int), ctx->stream));
    launch_filter(ctx, pred, B, N, nc, conf, legacy, classes, n_classes, cand, cand_count, cap);
    HBP_LAUNCH_CHECK(ctx);
    const int n_cap = std::min(cap, max_nms);
    const int words_all = (n_cap + 31) / 32;
    const size_t worst = (size_t)B * n_cap * words_all * sizeof(uint32_t);
    if (worst <= ((size_t)256 << 20)) {
        // the mask of the worst case (every row a candidate) fits a scratch buffer: no host round trip at all, the grids
        // are sized for the number of candidates a detector head realistically yields and stride beyond it
        uint32_t* mask = (uint32_t*)hbp_scratch(ctx, SC_NMS_MASK, worst);
        if (!mask) return HBP_ERR_NOMEM;
        return launch_nms_tail(ctx, cand, cand_count, sorted, n_sorted, mask, keep, keep_count, B, cap, std::min(n_cap, 2048), words_all,
                               (size_t)n_cap * words_all, thr, legacy, max_nms, max_keep, nullptr, out_det, out_count);
    }
    // large batches: the candidate count decides the mask size; read it back (4*B bytes) so the
    // mask scratch is sized for the actual n instead of N^2/8 bytes.
    int* h_counts = (int*)hbp_pinned(ctx, (size_t)B * sizeof(int));
    if (!h_counts) return HBP_ERR_NOMEM;
    HBP_CUDA(cudaMemcpyAsync(h_counts, cand_count, (size_t)B * sizeof(int), cudaMemcpyDeviceToHost, ctx->stream));
    HBP_CUDA(cudaStreamSynchronize(ctx->stream));
    int n_max = 0;
    for (int b = 0; b < B; ++b) n_max = max(n_max, min(h_counts[b], cap));
    n_max = min(n_max, max_nms);
    const int words = (std::max(n_max, 1) + 31) / 32;
    const size_t mask_img_stride = (size_t)std::max(n_max, 1) * words;
    uint32_t* mask = (uint32_t*)hbp_scratch(ctx, SC_NMS_MASK, (size_t)B * mask_img_stride * sizeof(uint32_t));
    if (!mask) return HBP_ERR_NOMEM;
    return launch_nms_tail(ctx, cand, cand_count, sorted, n_sorted, mask, keep, keep_count, B, cap, std::max(n_max, 1), words, mask_img_stride, thr,
                           legacy, max_nms, max_keep, nullptr, out_det, out_count);
}

// The same kernels with a fixed candidate capacity (the chained det -> pose pipeline): status bit 0 is set when an image had
// more candidates than cand_cap (the surplus was dropped in arrival order).
int run_nms_bounded(hbp_ctx* ctx, const float* pred, int B, int N, int nc, float conf, double thr,
                    const int* classes, int n_classes, int max_keep, int cand_cap, float* out_det, int* out_count,
                    int* status) {
    if (cand_cap > kMaxWords * 32 || cand_cap % 32) { hbp_set_error("candidate capacity must be a multiple of 32, at most %d", kMaxWords * 32); return HBP_ERR_INVALID; }
    const int cap = cand_cap;
    Cand* cand = (Cand*)hbp_scratch(ctx, SC_NMS_CAND, (size_t)B * cap * sizeof(Cand));
    SortedBox* sorted = (SortedBox*)hbp_scratch(ctx, SC_NMS_SORTED, (size_t)B * cap * sizeof(SortedBox));
    int* misc = (int*)hbp_scratch(ctx, SC_NMS_MISC, ((size_t)3 * B + (size_t)B * max_keep) * sizeof(int));
    const int words = cap / 32;
    const size_t mask_img_stride = (size_t)cap * words;
    uint32_t* mask = (uint32_t*)hbp_scratch(ctx, SC_NMS_MASK, (size_t)B * mask_img_stride * sizeof(uint32_t));
    if (!cand || !sorted || !misc || !mask) return HBP_ERR_NOMEM;
    int* cand_count = misc, *n_sorted = misc + B, *keep_count = misc + 2 * B, *keep = misc + 3 * B;
    HBP_CUDA(cudaMemsetAsync(misc, 0, (size_t)3 * B * sizeof(int), ctx->stream));
    launch_filter(ctx, pred, B, N, nc, conf, 0, classes, n_classes, cand, cand_count, cap);
    HBP_LAUNCH_CHECK(ctx);
    return launch_nms_tail(ctx, cand, cand_count, sorted, n_sorted, mask, keep, keep_count, B, cap, cap, words, mask_img_stride, thr, 0, 30000,
                           max_keep, status, out_det, out_count);
}

// ---- detections -> per-person crop parameters, on the device --------------------------------------------------
// One CTA.  Persons are numbered frame by frame in detector order; person p >= persons_cap is dropped (status bit 1).
// Outputs per person: 2x3 dst->src matrix (double, what hbp_crop_warp takes), the pixel box yxyx (float, what
// hbp_decode_proportions takes), its frame and its height_cm = heights[min(i, n_heights-1)], i = index inside the
// frame (person_det_pose_edet4_trtserver.py:166-168).
//
// YOLO (configs[2]): det rows [x1,y1,x2,y2,conf,cls] in letterbox pixels -> scale_coords + clip_coords
// (onnx_utils.py:238-266, the arithmetic of scale_coords_kernel) -> int() truncation, the crop the reference's
// PoseEstimator.preprocess would take from frame[y1:y2, x1:x2] with cv2.resize (pose_estimator.py:29-45): the
// half-pixel stretch matrix of geometry.box_resize_matrices.
__global__ void __launch_bounds__(256)
persons_from_yolo_kernel(const float* __restrict__ det, const int* __restrict__ det_count, int F, int max_det,
                         float pad_x, float pad_y, float gain, float w0, float h0, int out_h, int out_w,
                         const double* __restrict__ heights, int n_heights, int persons_cap,
                         double* __restrict__ M, float* __restrict__ boxes, int* __restrict__ frame_idx,
                         double* __restrict__ height_cm, int* __restrict__ n_persons, int* __restrict__ status) {
    __shared__ int s_off[1025];
    if (threadIdx.x == 0) {
        int o = 0;
        for (int f = 0; f < F; ++f) { s_off[f] = o; o += min(det_count[f], max_det); }
        s_off[F] = o;
        *n_persons = min(o, persons_cap);
        if (o > persons_cap) atomicOr(status, 2);
    }
    __syncthreads();
    for (int f = 0; f < F; ++f) {
        const int n = s_off[f + 1] - s_off[f];
        for (int i = threadIdx.x; i < n; i += blockDim.x) {
            const int p = s_off[f] + i;
            if (p >= persons_cap) continue;
            const float* b = det + ((size_t)f * max_det + i) * 6;
            float x1 = __fdiv_rn(__fsub_rn(b[0], pad_x), gain), y1 = __fdiv_rn(__fsub_rn(b[1], pad_y), gain);
            float x2 = __fdiv_rn(__fsub_rn(b[2], pad_x), gain), y2 = __fdiv_rn(__fsub_rn(b[3], pad_y), gain);
            x1 = fminf(fmaxf(x1, 0.f), w0); y1 = fminf(fmaxf(y1, 0.f), h0);
            x2 = fminf(fmaxf(x2, 0.f), w0); y2 = fminf(fmaxf(y2, 0.f), h0);
            const int xi1 = (int)x1, yi1 = (int)y1, xi2 = (int)x2, yi2 = (int)y2;
            const double sx = (double)(xi2 - xi1) / (double)out_w, sy = (double)(yi2 - yi1) / (double)out_h;
            double* m = M + (size_t)p * 6;
            m[0] = sx; m[1] = 0.0; m[2] = (double)xi1 + 0.5 * sx - 0.5;
            m[3] = 0.0; m[4] = sy; m[5] = (double)yi1 + 0.5 * sy - 0.5;
            float* o = boxes + (size_t)p * 4;
            o[0] = (float)yi1; o[1] = (float)xi1; o[2] = (float)yi2; o[3] = (float)xi2;
            frame_idx[p] = f;
            height_cm[p] = heights[min(i, n_heights - 1)];
        }
    }
}

// EfficientDet (configs[3]): filtered boxes (F, max_persons, 4) yxyx NORMALISED + counts -> tf.image.crop_and_resize
// matrices (models/conv.py:61-70; geometry.crop_and_resize_matrices in double) and boxes * [h,w,h,w] in float32
// (person_det_pose_edet4_trtserver.py:145).
__global__ void __launch_bounds__(256)
persons_from_edet_kernel(const float* __restrict__ boxes_n, const int* __restrict__ counts, int F, int max_persons,
                         int img_h, int img_w, int out_h, int out_w, const double* __restrict__ heights, int n_heights,
                         int persons_cap, double* __restrict__ M, float* __restrict__ boxes, int* __restrict__ frame_idx,
                         double* __restrict__ height_cm, int* __restrict__ n_persons, int* __restrict__ status) {
    __shared__ int s_off[1025];
    if (threadIdx.x == 0) {
        int o = 0;
        for (int f = 0; f < F; ++f) { s_off[f] = o; o += min(counts[f], max_persons); }
        s_off[F] = o;
        *n_persons = min(o, persons_cap);
        if (o > persons_cap) atomicOr(status, 2);
    }
    __syncthreads();
    const double dw = (double)(out_w - 1 > 1 ? out_w - 1 : 1), dh = (double)(out_h - 1 > 1 ? out_h - 1 : 1);
    for (int t = threadIdx.x; t < F * max_persons; t += blockDim.x) {
        const int f = t / max_persons, i = t % max_persons;
        if (i >= s_off[f + 1] - s_off[f]) continue;
        const int p = s_off[f] + i;
        if (p >= persons_cap) continue;
        const float* b = boxes_n + (size_t)t * 4;
        const double y1 = b[0], x1 = b[1], y2 = b[2], x2 = b[3];
        double* m = M + (size_t)p * 6;
        m[0] = __ddiv_rn(__dmul_rn(__dsub_rn(x2, x1), (double)(img_w - 1)), dw); m[1] = 0.0; m[2] = __dmul_rn(x1, (double)(img_w - 1));
        m[3] = 0.0; m[4] = __ddiv_rn(__dmul_rn(__dsub_rn(y2, y1), (double)(img_h - 1)), dh); m[5] = __dmul_rn(y1, (double)(img_h - 1));
        float* o = boxes + (size_t)p * 4;
        o[0] = __fmul_rn(b[0], (float)img_h); o[1] = __fmul_rn(b[1], (float)img_w);
        o[2] = __fmul_rn(b[2], (float)img_h); o[3] = __fmul_rn(b[3], (float)img_w);
        frame_idx[p] = f;
        height_cm[p] = heights[min(i, n_heights - 1)];
    }
}

}  // namespace

int k_yolo_filter(hbp_ctx* ctx, const float* pred, int B, int N, int nc, float conf, const int* classes, int n_classes,
                  int cand_cap, int* out_count) {
    Cand* cand = (Cand*)hbp_scratch(ctx, SC_NMS_CAND, (size_t)B * cand_cap * sizeof(Cand));
    if (!cand) return HBP_ERR_NOMEM;
    HBP_CUDA(cudaMemsetAsync(out_count, 0, (size_t)B * sizeof(int), ctx->stream));
    launch_filter(ctx, pred, B, N, nc, conf, 0, classes, n_classes, cand, out_count, cand_cap);
    HBP_LAUNCH_CHECK(ctx);
    return HBP_OK;
}

int k_yolo_nms_bounded(hbp_ctx* ctx, const float* pred, int B, int N, int nc, float conf, double iou,
                       const int* classes, int n_classes, int max_det, int cand_cap, float* out_det, int* out_count,
                       int* status) {
    return run_nms_bounded(ctx, pred, B, N, nc, conf, iou, classes, n_classes, max_det, cand_cap, out_det, out_count, status);
}

int k_persons_from_yolo(hbp_ctx* ctx, const float* det, const int* det_count, int F, int max_det, int in_h, int in_w,
                        int img_h, int img_w, int out_h, int out_w, const double* heights, int n_heights,
                        int persons_cap, double* M, float* boxes, int* frame_idx, double* height_cm,
                        int* n_persons, int* status) {
    if (F > 1024) { hbp_set_error("at most 1024 frames per call"); return HBP_ERR_INVALID; }
    const double gain = (double)(in_h > in_w ? in_h : in_w) / (double)(img_h > img_w ? img_h : img_w);
    const double pad_x = ((double)in_w - (double)img_w * gain) / 2.0, pad_y = ((double)in_h - (double)img_h * gain) / 2.0;
    persons_from_yolo_kernel<<<1, 256, 0, ctx->stream>>>(det, det_count, F, max_det, (float)pad_x, (float)pad_y, (float)gain,
                                                        (float)img_w, (float)img_h, out_h, out_w, heights, n_heights,
                                                        persons_cap, M, boxes, frame_idx, height_cm, n_persons, status);
    HBP_LAUNCH_CHECK(ctx);
    return HBP_OK;
}

int k_persons_from_edet(hbp_ctx* ctx, const float* boxes_n, const int* counts, int F, int max_persons, int img_h,
                        int img_w, int out_h, int out_w, const double* heights, int n_heights, int persons_cap,
                        double* M, float* boxes, int* frame_idx, double* height_cm, int* n_persons, int* status) {
    if (F > 1024) { hbp_set_error("at most 1024 frames per call"); return HBP_ERR_INVALID; }
    persons_from_edet_kernel<<<1, 256, 0, ctx->stream>>>(boxes_n, counts, F, max_persons, img_h, img_w, out_h, out_w,
                                                        heights, n_heights, persons_cap, M, boxes, frame_idx, height_cm,
                                                        n_persons, status);
    HBP_LAUNCH_CHECK(ctx);
    return HBP_OK;
}

int k_yolo_decode_raw(hbp_ctx* ctx, const float* h0, const float* h1, const float* h2, int B, int s0,
                      int s1, int s2, int nc, int in_w, int in_h, float* out) {
    const int E = 5 + nc;
    const float* heads[3] = {h0, h1, h2};
    const int S[3] = {s0, s1, s2};
    const int total_rows = 3 * (s0 * s0 + s1 * s1 + s2 * s2);
    int row_off = 0;
    for (int l = 0; l < 3; ++l) {
        // obj_det_yolov5_onnx.py:145-146: stride = int(in / feature)
        const float sw = (float)(int)((double)in_w / S[l]), sh = (float)(int)((double)in_h / S[l]);
        const size_t total = (size_t)B * 3 * S[l] * S[l] * E;
        const int blocks = (int)std::min<size_t>((total + 255) / 256, (size_t)ctx->sm_count * 16);
        yolo_decode_kernel<<<blocks, 256, 0, ctx->stream>>>(heads[l], out, B, S[l], E, l, sw, sh, row_off, total_rows);
        HBP_LAUNCH_CHECK(ctx);
        row_off += 3 * S[l] * S[l];
    }
    return HBP_OK;
}

int k_yolo_nms(hbp_ctx* ctx, const float* pred, int B, int N, int nc, float conf, double iou,
               const int* classes, int n_classes, int max_det, float* out_det, int* out_count) {
    return run_nms(ctx, pred, B, N, nc, conf, iou, classes, n_classes, max_det, 0, out_det, out_count);
}

int k_yolo_nms_legacy(hbp_ctx* ctx, const float* pred, int B, int N, int nc, float conf, float thr,
                      int max_out, float* out_det, int* out_count) {
    return run_nms(ctx, pred, B, N, nc, conf, (double)thr, nullptr, 0, max_out, 1, out_det, out_count);
}

int k_scale_coords(hbp_ctx* ctx, float* boxes, int n, int h1, int w1, int h0, int w0) {
    // python: gain = max(img1)/max(img0); pad = ((w1 - w0*gain)/2, (h1 - h0*gain)/2) in double,
    // then applied to float32 coords
    const double gain = (double)(h1 > w1 ? h1 : w1) / (double)(h0 > w0 ? h0 : w0);
    const double pad_x = ((double)w1 - (double)w0 * gain) / 2.0, pad_y = ((double)h1 - (double)h0 * gain) / 2.0;
    scale_coords_kernel<<<(n + 127) / 128, 128, 0, ctx->stream>>>(boxes, n, (float)pad_x, (float)pad_y, (float)gain, (float)w0, (float)h0);
    HBP_LAUNCH_CHECK(ctx);
    return HBP_OK;
}

int k_edet_filter(hbp_ctx* ctx, const float* boxes, const float* scores, const float* classes, int F,
                  int K, float person_class, float thr, float xe, float ye, int img_h, int img_w,
                  int max_persons, float* out_boxes, int* out_count) {
    HBP_CUDA(cudaMemsetAsync(out_boxes, 0, (size_t)F * max_persons * 4 * sizeof(float), ctx->stream));
    edet_filter_kernel<<<F, 32, 0, ctx->stream>>>(boxes, scores, classes, K, person_class, thr, xe, ye,
                                                  (float)img_h, (float)img_w, max_persons, out_boxes, out_count);
    HBP_LAUNCH_CHECK(ctx);
    return HBP_OK;
}
