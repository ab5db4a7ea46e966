// K5 engine 1: convolution as implicit GEMM on the 5th-generation tensor cores.
//
//   D[pixels, Cout] (fp32, TMEM)  +=  A[pixels, (tap, cin)] (fp16, smem)  x  B[Cout, (tap, cin)]^T (fp16, smem)
//
// * A is never materialised: for every filter tap the TMA engine loads the
//   shifted (tn x th x tw) pixel box of the NHWC activation tensor -- a 4-D
//   tiled tensor map (C, W, H, N) whose out-of-bounds zero fill IS the
//   convolution's zero padding; stride-2 convolutions use the map's element
//   strides.  The box lands in shared memory as 128-byte (64-channel) or
//   64-byte (32-channel) rows in the hardware swizzle, which is exactly the
//   K-major canonical layout tcgen05.mma reads through a shared-memory
//   descriptor.  B (weights, [tap][cout][cin]) comes through a 2-D map.
// * One CTA = 128*m_tiles output pixels x n_tile output channels.  Warp 0
//   produces (TMA + mbarrier expect_tx), lane 0 of warp 1 issues
//   tcgen05.mma.cta_group::1.kind::f16 (M=128, N=n_tile, K=16) into TMEM and
//   commits to the stage's "empty" barrier, warps 2..5 are the epilogue:
//   tcgen05.ld 32x32b -> +bias (+residual) -> ReLU -> fp16 -> NHWC store, with
//   the fuse layers' nearest upsample folded into the store.
// * Several CTAs share an SM (small stages, <= 512 TMEM columns in total), so
//   one CTA's epilogue overlaps another's main loop.
//
// Bound: tensor pipe (2*pixels*Cout*Cin*k*k flop per launch); see DESIGN.md.
#include "hrnet.cuh"
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <algorithm>
#include <cmath>

// ---------------------------------------------------------------------------
// driver entry point for tensor-map encoding (no link-time libcuda dependency)
// ---------------------------------------------------------------------------
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*,
                                  CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion,
                                  CUtensorMapFloatOOBfill);

static EncodeTiledFn get_encode() {
    static EncodeTiledFn fn = nullptr;
    if (!fn) {
        void* p = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess &&
            q == cudaDriverEntryPointSuccess)
            fn = reinterpret_cast<EncodeTiledFn>(p);
    }
    return fn;
}

struct ConvParams {
    int P, Ho, Wo, Cout, up, relu;
    int tn, th, tw, m_tiles, n_tile;
    int chunk, n_chunks, ksz, stride;
    int tiles_w, tiles_h;
    int stages;
    uint32_t a_stage_bytes, b_stage_bytes, tx_bytes;
    uint32_t row_bytes;          // 64 or 128
    uint32_t tmem_cols;
    uint32_t idesc;
    // halo mode (3x3 stride 1): the (rs x 10)-pixel halo tile of every Cin chunk is
    // loaded ONCE and the nine taps are nine descriptor start offsets into it
    int mode;                    // 0 = one TMA box per tap, 1 = halo tile + shifted descriptors
    int rs;                      // stacked rows per image in the tile (th + 2)
    uint32_t a_chunk_bytes;      // smem reserved per Cin chunk (>= box, covers (16*m_tiles+2)*10 rows)
    uint32_t a_box_bytes;        // bytes one halo box writes
    const float* bias;
    const __half* res;
    __half* out;
    int n_tiles;                 // halo mode: tiles walked by the persistent CTAs (set at launch)
    int acc_bufs;                // halo mode: accumulator buffers in TMEM (always 2)
    int a_stages;                // halo mode: halo tiles in flight (ring depth)
    int b_slots;                 // halo mode: weight slots of three taps (one dx, dy = 0..2) each: 3*n_chunks when resident, else ring depth
    int b_resident;              // halo mode: weights stay in shared memory for the CTA's lifetime
    int res_smem;                // halo mode: residual tiles come through a TMA ring in shared memory (same stage index as the halo tile)
    int r_chunks;                // residual boxes per tile (64 output channels each)
    uint32_t r_chunk_bytes, r_box_bytes, r_row_bytes;
    int issuers;                 // halo mode: MMA-issuing warps
    int teams;                   // halo mode: epilogue teams (2 needs an even number of ring stages and accumulator buffers) (2: alternate tiles, one accumulator buffer each)
    int phase_maps;              // stride-2 per-tap mode: the four (row, column) parity planes of the input have their own dense tensor maps
    int b_sets;                  // chain kernel, resident weights: 1 = one set reloaded at every layer switch, 2 = the next layer's set is prefetched
    int bias_mma;                // halo mode: the bias enters the accumulator through one extra MMA per M-tile (ones x [bias_hi, bias_lo]) instead of the epilogue
    int res_mma;                 // halo mode: the residual tile (TMA ring) is added by MMAs against a 32x32 identity instead of the epilogue
    uint32_t c_bytes;            // halo mode: constant operand tiles (ones, identity, bias) between the residual ring and the barriers
    uint32_t idesc32;            // instruction descriptor with N = 32 (residual MMAs)
    int reverse;                 // halo mode: the CTAs walk the tiles from the last to the first (tile = n_tiles - 1 - index)
    int split_producer;          // halo mode, streamed weights: halo tiles are requested by warp 3 chunk by chunk, warp 0 only streams weights
    unsigned long long hint_a, hint_r;   // halo mode: L2 cache policies of the halo / residual TMA loads (kEvictNormal, kEvictFirst)
    unsigned long long hint_o;           // halo mode: L2 cache policy of the output stores (0 = plain stores)
    int dbg_flags;               // bring-up experiments: 1 = skip the output stores, 2 = skip the bias loads
    long long* dbg;              // optional phase timestamps (8 per CTA, first 64 CTAs), bring-up only
    unsigned long long* tl;      // optional [start, end] globaltimer stamps of this launch (HBP_TIMELINE)
};

struct PhaseMaps { CUtensorMap m[4]; };     // [row parity * 2 + column parity]

struct UmmaPlan {
    CUtensorMap tmA, tmB, tmR;   // tmR: residual tensor (halo mode with res_smem)
    PhaseMaps tmP;               // stride-2 per-tap convs (prm.phase_maps)
    ConvParams prm;
    size_t smem_bytes;
    int n_splits;
    int occ;                     // resident CTAs per SM (halo mode: sizes the persistent grid)
    int sm_budget;               // halo mode: SMs (= persistent CTAs) this launch may use
};

// ---------------------------------------------------------------------------
// PTX wrappers
// ---------------------------------------------------------------------------
namespace {

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
// Watchdog of every in-kernel wait (mbarrier phases, the chain kernel's dependency flags): SM clock cycles after which a
// wait that has not advanced traps instead of hanging the GPU (a wrong expect_tx byte count or tensor map never
// completes).  clock64 keeps counting while a context is time-sliced or replayed by a profiler, so the limit is generous
// (HBP_WATCHDOG_S seconds at ~2 GHz, default 20; 0 disables it) -- set once per process by umma_plan_create.
__device__ long long g_watchdog_cycles = 40000000000LL;

__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
    // try_wait suspends the warp in hardware up to the time hint (it wakes on completion), so a
    // waiting role does not take issue slots from the warps that share its scheduler
    long long t0 = 0;
    for (uint32_t spins = 0;; ++spins) {
        uint32_t done;
        asm volatile(
            "{\n"
            ".reg .pred p;\n"
            "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2, %3;\n"
            "selp.u32 %0, 1, 0, p;\n"
            "}\n" : "=r"(done) : "r"(bar), "r"(parity), "r"(0x989680u) : "memory");
        if (done) return;
        if ((spins & 1023u) == 1023u) {
            const long long now = clock64();
            if (t0 == 0) t0 = now;
            else if (g_watchdog_cycles > 0 && now - t0 > g_watchdog_cycles) __trap();
        }
    }
}
__device__ __forceinline__ void tma_load_4d(uint32_t dst, const CUtensorMap* map, uint32_t bar, int c0, int c1, int c2, int c3) {
    asm volatile(
        "cp.async.bulk.tensor.4d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2];"
        ::"r"(dst), "l"(map), "r"(bar), "r"(c0), "r"(c1), "r"(c2), "r"(c3) : "memory");
}
// L2 eviction policies in the encoding `createpolicy` produces (what CUTLASS passes as TMA::CacheHintSm90)
constexpr unsigned long long kEvictNormal = 0x1000000000000000ull, kEvictFirst = 0x12F0000000000000ull, kEvictLast = 0x14F0000000000000ull;
__device__ __forceinline__ void tma_load_4d_hint(uint32_t dst, const CUtensorMap* map, uint32_t bar, int c0, int c1, int c2, int c3,
                                                 unsigned long long policy) {
    asm volatile(
        "cp.async.bulk.tensor.4d.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint [%0], [%1, {%3, %4, %5, %6}], [%2], %7;"
        ::"r"(dst), "l"(map), "r"(bar), "r"(c0), "r"(c1), "r"(c2), "r"(c3), "l"(policy) : "memory");
}
// 16 output channels of one pixel = one full 32-byte sector: a single STG.256 when the address allows it (always, for
// channel counts that are multiples of 16), else two 16-byte stores
__device__ __forceinline__ void store_half16(__half* dst, const __half2* pk) {
    const uint4 v0 = *reinterpret_cast<const uint4*>(pk), v1 = *reinterpret_cast<const uint4*>(pk + 4);
    if ((reinterpret_cast<uintptr_t>(dst) & 31u) == 0) {
        asm volatile("st.global.v8.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};" ::"l"(dst), "r"(v0.x), "r"(v0.y), "r"(v0.z), "r"(v0.w),
                     "r"(v1.x), "r"(v1.y), "r"(v1.z), "r"(v1.w) : "memory");
    } else {
        *reinterpret_cast<uint4*>(dst) = v0;
        *reinterpret_cast<uint4*>(dst + 8) = v1;
    }
}
// persistent tile walk of a halo CTA: tile index -> (column tile, row tile, image group) advanced by the grid size without
// divisions, forwards (tile = blockIdx.x + j * grid) or backwards (tile = n_tiles - 1 - blockIdx.x - j * grid)
struct TileWalk {
    int tw, th, tg, dw, dh, dg, tiles_w, tiles_h, rev;
    __device__ __forceinline__ void init(int first, int step, int tiles_w_, int tiles_h_, int n_tiles, int reverse) {
        tiles_w = tiles_w_; tiles_h = tiles_h_; rev = reverse;
        const int per_img = tiles_w * tiles_h;
        const int tile = reverse ? n_tiles - 1 - first : first;
        tw = tile % tiles_w; th = (tile / tiles_w) % tiles_h; tg = tile / per_img;
        dw = step % tiles_w; dh = (step / tiles_w) % tiles_h; dg = step / per_img;
    }
    __device__ __forceinline__ void next() {
        if (!rev) {
            tw += dw; if (tw >= tiles_w) { tw -= tiles_w; ++th; }
            th += dh; if (th >= tiles_h) { th -= tiles_h; ++tg; }
            tg += dg;
        } else {
            tw -= dw; if (tw < 0) { tw += tiles_w; --th; }
            th -= dh; if (th < 0) { th += tiles_h; --tg; }
            tg -= dg;
        }
    }
};
__device__ __forceinline__ void tma_load_2d(uint32_t dst, const CUtensorMap* map, uint32_t bar, int c0, int c1) {
    asm volatile(
        "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
        ::"r"(dst), "l"(map), "r"(bar), "r"(c0), "r"(c1) : "memory");
}
__device__ __forceinline__ void tmem_alloc(uint32_t dst_smem, uint32_t cols) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(dst_smem), "r"(cols) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr, uint32_t cols) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(cols) : "memory");
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void umma_commit(uint32_t bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void umma_f16(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "setp.ne.b32 p, %4, 0;\n"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n"
        "}\n" ::"r"(d_tmem), "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(accumulate) : "memory");
}
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, uint32_t (&r)[16]) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
          "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
        : "r"(taddr) : "memory");
}
// One lane of a converged warp.  tcgen05.mma / TMA / commit take their operands from the
// warp-uniform register file: the role loops below therefore run warp-wide on provably
// uniform values (kernel parameters, loop counters, shuffled warp index) and only the issue
// itself sits under elect.sync -- inside a divergent `if (lane == 0)` the compiler has to wrap
// every such instruction in an ELECT / R2UR.BROADCAST / BRA.U.ANY waterfall loop (~200 cycles
// per MMA, measured: profiles/r01_conv_phase_trace.log).
__device__ __forceinline__ bool elect_one() {
    uint32_t pred;
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "elect.sync _|p, 0xffffffff;\n"
        "selp.u32 %0, 1, 0, p;\n"
        "}\n" : "=r"(pred));
    return pred != 0;
}
__device__ __forceinline__ int uniform_warp_idx() { return __shfl_sync(0xffffffffu, (int)(threadIdx.x >> 5), 0); }

__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

// shared-memory matrix descriptor, K-major, hardware swizzle (SM100 format):
//   [0,14) start>>4 | [16,30) LBO>>4 (unused for swizzled K-major, 1) | [32,46) SBO>>4
//   [46,48) version=1 | [61,64) layout: 2 = SWIZZLE_128B, 4 = SWIZZLE_64B
//   base_offset stays 0: tcgen05.mma (like TMA) swizzles on ABSOLUTE shared-memory address
//   bits, so start addresses and group strides need not be pattern-aligned
//   (measured: tools/umma_probe.cu, profiles/r01_umma_swizzle_probe.log).
__device__ __forceinline__ uint64_t make_desc(uint32_t saddr, uint32_t row_bytes, uint32_t sbo_bytes = 0) {
    uint64_t d = 0;
    d |= (uint64_t)((saddr & 0x3FFFF) >> 4);
    d |= (uint64_t)1 << 16;
    d |= (uint64_t)((sbo_bytes ? sbo_bytes : 8u * row_bytes) >> 4) << 32;
    d |= (uint64_t)1 << 46;
    d |= (uint64_t)(row_bytes == 128 ? 2 : 4) << 61;
    return d;
}

// Programmatic dependent launch: a kernel launched with the programmatic-serialization attribute
// starts while its stream predecessor is still running; everything that does not touch the
// predecessor's output (barrier init, TMEM allocation, weight / bias loads) runs ahead of
// pdl_wait(), which returns once the predecessor grid has completed and flushed.
__device__ __forceinline__ unsigned long long globaltimer_ns() {
    unsigned long long t;
    asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t));
    return t;
}
__device__ __forceinline__ void pdl_launch_dependents() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }

constexpr int kThreads = 192;    // warp 0 TMA, warp 1 MMA + TMEM alloc, warps 2-5 epilogue

// One output tile of one convolution (mode 0: one TMA box per tap).  Shared by the single-conv
// kernel and the grouped kernel below; `bx` / `by` are the tile / output-channel-split indices.
__device__ __forceinline__ void conv_umma_body(const CUtensorMap* tmAp, const CUtensorMap* tmBp, const CUtensorMap* tmPp, const ConvParams& p,
                                               const int bx, const int by, uint8_t* smem_raw) {
    // carve: [A stages][B stages] (1024-aligned), then barriers
    const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
    const uint32_t a_base = smem_base;
    const uint32_t b_base = a_base + p.stages * p.a_stage_bytes;
    const uint32_t bar_base = b_base + p.stages * p.b_stage_bytes;      // 8-byte aligned
    const uint32_t full_bar = bar_base;                                 // stages x 8 B
    const uint32_t empty_bar = bar_base + 8u * p.stages;
    const uint32_t tmem_full_bar = bar_base + 16u * p.stages;
    const uint32_t tmem_slot = tmem_full_bar + 8u;
    float* s_bias = reinterpret_cast<float*>(smem_raw + (tmem_slot + 8u - smem_u32(smem_raw)));

    const int warp = uniform_warp_idx(), lane = threadIdx.x & 31;
    const int taps = p.ksz * p.ksz;
    const int k_iters = taps * p.n_chunks;

    // tile coordinates
    int t = bx;
    const int tile_w = t % p.tiles_w; t /= p.tiles_w;
    const int tile_h = t % p.tiles_h; t /= p.tiles_h;
    const int n0 = t * p.tn, h0 = tile_h * p.th, w0 = tile_w * p.tw;
    const int n_off = by * p.n_tile;

    pdl_launch_dependents();             // the next launch of this stream may start its prologue
    if (p.tl && threadIdx.x == 0) atomicMin(p.tl, globaltimer_ns());
    if (warp == 0) {
        if (lane == 0) {
            asm volatile("prefetch.tensormap [%0];" ::"l"(tmAp) : "memory");
            asm volatile("prefetch.tensormap [%0];" ::"l"(tmBp) : "memory");
            mbar_init(tmem_full_bar, 1);
        }
        for (int s = lane; s < p.stages; s += 32) {
            mbar_init(full_bar + 8u * s, 1);
            mbar_init(empty_bar + 8u * s, 1);
        }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 1) tmem_alloc(tmem_slot, p.tmem_cols);
    for (int i = threadIdx.x; i < p.n_tile; i += kThreads) s_bias[i] = p.bias[by * p.n_tile + i];
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    pdl_wait();                          // activations / residual below are the previous launch's output
    uint32_t tmem_base;
    asm volatile("ld.shared.b32 %0, [%1];" : "=r"(tmem_base) : "r"(tmem_slot));
    tmem_base = __shfl_sync(0xffffffffu, tmem_base, 0);

    if (warp == 0) {
        // ===== TMA producer (warp-wide control flow, one elected lane issues) =====
        const int pad = p.ksz / 2;
        int s = 0;
        uint32_t ph = 0;
        // K order = (Cin chunk, dx, dy, 16-channel step), the order of the halo kernel: a conv's fp32 sums, and
        // with them the network's outputs, do not depend on which of the two kernels a plan picks
        for (int kk = 0; kk < taps * p.n_chunks; ++kk) {
            const int cc = kk / taps, tk = kk % taps;
            const int dxi = tk / p.ksz, dyi = tk % p.ksz, tap = dyi * p.ksz + dxi;
            const int dy = dyi - pad, dx = dxi - pad;
            {
                mbar_wait(empty_bar + 8u * s, ph ^ 1u);
                if (elect_one()) {
                    mbar_expect_tx(full_bar + 8u * s, p.tx_bytes);
                    if (p.phase_maps)          // tap (dy,dx) of a stride-2 conv = a dense box of one parity plane (see umma_plan_create)
                        tma_load_4d(a_base + s * p.a_stage_bytes, tmPp + ((dy & 1) * 2 + (dx & 1)), full_bar + 8u * s,
                                    cc * p.chunk, w0 - (dx < 0), h0 - (dy < 0), n0);
                    else
                        tma_load_4d(a_base + s * p.a_stage_bytes, tmAp, full_bar + 8u * s,
                                    cc * p.chunk, w0 * p.stride + dx, h0 * p.stride + dy, n0);
                    tma_load_2d(b_base + s * p.b_stage_bytes, tmBp, full_bar + 8u * s,
                                cc * p.chunk, tap * p.Cout + n_off);
                }
                __syncwarp();
                if (++s == p.stages) { s = 0; ph ^= 1u; }
            }
        }
    } else if (warp == 1) {
        // ===== MMA issuer (warp-wide control flow, one elected lane issues) =====
        {
            // descriptors differ only in their 14-bit start-address field (16-byte units):
            // build the constant part once, add offsets in the loop (the issuing thread is a
            // single lane, every instruction it spends is serial latency)
            const int ksteps = p.chunk / 16;
            const uint64_t d0 = make_desc(0, p.row_bytes);
            const uint32_t mt_step = (128u * p.row_bytes) >> 4;
            int s = 0;
            uint32_t ph = 0;
            for (int it = 0; it < k_iters; ++it) {
                mbar_wait(full_bar + 8u * s, ph);
                tc_fence_after();
                const uint64_t ad0 = d0 + ((a_base + s * p.a_stage_bytes) >> 4);
                const uint64_t bd0 = d0 + ((b_base + s * p.b_stage_bytes) >> 4);
                if (elect_one()) {
                    for (int mt = 0; mt < p.m_tiles; ++mt) {
                        const uint64_t ad = ad0 + mt * mt_step;
                        const uint32_t dt = tmem_base + mt * p.n_tile;
                        umma_f16(dt, ad, bd0, p.idesc, it ? 1u : 0u);
                        umma_f16(dt, ad + 2, bd0 + 2, p.idesc, 1u);
                        if (ksteps == 4) {
                            umma_f16(dt, ad + 4, bd0 + 4, p.idesc, 1u);
                            umma_f16(dt, ad + 6, bd0 + 6, p.idesc, 1u);
                        }
                    }
                    umma_commit(empty_bar + 8u * s);      // frees the stage when these MMAs retire
                }
                __syncwarp();
                if (++s == p.stages) { s = 0; ph ^= 1u; }
            }
            if (elect_one()) umma_commit(tmem_full_bar);
            __syncwarp();
        }
    } else {
        // ===== epilogue: TMEM -> registers -> (+bias, +residual, ReLU) -> global =====
        const int grp = warp & 3;                          // TMEM lane group this warp may read
        const int Hout = p.Ho * p.up, Wout = p.Wo * p.up;
        bool valid[2];
        int pn[2], ph_[2], pw[2];
#pragma unroll
        for (int mt = 0; mt < 2; ++mt) {
            const int R = mt * 128 + grp * 32 + lane;
            pw[mt] = w0 + R % p.tw; ph_[mt] = h0 + (R / p.tw) % p.th; pn[mt] = n0 + R / (p.tw * p.th);
            valid[mt] = mt < p.m_tiles && pn[mt] < p.P;
        }
        // residual of the first 64 columns of the first M-tile is in flight before the accumulator is ready
        const bool direct = p.up == 1;
        uint4 rq[8];
        auto fetch_res = [&](int mt, int cg) {
            const size_t o = ((((size_t)pn[mt] * Hout + ph_[mt]) * Wout) + pw[mt]) * p.Cout + n_off + cg;
#pragma unroll
            for (int q = 0; q < 8; ++q)
                if (cg + q * 8 < p.n_tile) rq[q] = *reinterpret_cast<const uint4*>(p.res + o + q * 8);
        };
        if (direct && p.res && valid[0]) fetch_res(0, 0);
        mbar_wait(tmem_full_bar, 0);
        tc_fence_after();
        for (int mt = 0; mt < p.m_tiles; ++mt) {
            for (int cg = 0; cg < p.n_tile; cg += 64) {
                if (direct && p.res && valid[mt] && (mt | cg)) fetch_res(mt, cg);
#pragma unroll
                for (int cq = 0; cq < 4; ++cq) {
                    const int c0 = cg + cq * 16;
                    if (c0 >= p.n_tile) break;
                    uint32_t r[16];
                    tmem_ld16(tmem_base + ((uint32_t)(grp * 32) << 16) + (uint32_t)(mt * p.n_tile + c0), r);
                    tmem_ld_wait();
                    if (!valid[mt]) continue;
                    float v[16];
#pragma unroll
                    for (int q = 0; q < 16; ++q) v[q] = __uint_as_float(r[q]) + s_bias[c0 + q];
                    for (int uy = 0; uy < p.up; ++uy)
                        for (int ux = 0; ux < p.up; ++ux) {
                            const size_t o = ((((size_t)pn[mt] * Hout + ph_[mt] * p.up + uy) * Wout) + pw[mt] * p.up + ux) * p.Cout + n_off + c0;
                            float x[16];
#pragma unroll
                            for (int q = 0; q < 16; ++q) x[q] = v[q];
                            if (p.res) {
                                uint4 q0, q1;
                                if (direct) { q0 = rq[2 * cq]; q1 = rq[2 * cq + 1]; }
                                else {
                                    q0 = *reinterpret_cast<const uint4*>(p.res + o);
                                    q1 = *reinterpret_cast<const uint4*>(p.res + o + 8);
                                }
                                const __half2* h0p = reinterpret_cast<const __half2*>(&q0);
                                const __half2* h1p = reinterpret_cast<const __half2*>(&q1);
#pragma unroll
                                for (int q = 0; q < 4; ++q) {
                                    const float2 f0 = __half22float2(h0p[q]), f1 = __half22float2(h1p[q]);
                                    x[2 * q] += f0.x; x[2 * q + 1] += f0.y;
                                    x[8 + 2 * q] += f1.x; x[8 + 2 * q + 1] += f1.y;
                                }
                            }
                            if (p.relu) {
#pragma unroll
                                for (int q = 0; q < 16; ++q) x[q] = fmaxf(x[q], 0.f);
                            }
                            __align__(16) __half2 pk[8];
#pragma unroll
                            for (int q = 0; q < 8; ++q) pk[q] = __floats2half2_rn(x[2 * q], x[2 * q + 1]);
                            store_half16(p.out + o, pk);
                        }
                }
            }
        }
        tc_fence_before();
    }
    __syncthreads();
    if (warp == 1) {
        tc_fence_after();
        tmem_dealloc(tmem_base, p.tmem_cols);
    }
    if (p.tl && threadIdx.x == 0) atomicMax(p.tl + 1, globaltimer_ns());
}


__global__ void __launch_bounds__(kThreads)
conv_umma_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
                 const __grid_constant__ PhaseMaps tmP, const ConvParams p) {
    extern __shared__ __align__(1024) uint8_t smem_raw[];
    conv_umma_body(&tmA, &tmB, &tmP.m[0], p, p.reverse ? (int)(gridDim.x - 1u - blockIdx.x) : (int)blockIdx.x, (int)blockIdx.y, smem_raw);
}

// Persistent grouped launch (mode 0: one TMA box per tap).  A group is a list of independent
// convolutions -- the links of the stride-2 chains and the low-resolution 1x1 convs of one fuse level,
// or a single stride-2 convolution -- whose (conv, tile, output-channel split) work items form one
// list that <= one CTA per SM walks with a stride.  A one-tile-per-CTA grid spends ~10 us of
// TMEM allocation / barrier setup / pipeline fill / drain on ~0.5 us of MMAs; here the prologue is
// paid once per SM, the TMA ring runs on across item boundaries, and the accumulator is double-buffered
// in TMEM so that the epilogue of item k overlaps the MMAs of item k+1.  The member plans (tensor
// maps + parameters) sit in a device table, item ranges in the by-value header.
constexpr int kMaxGroup = 16;
constexpr int kMaxPStages = 12;
struct alignas(64) GroupEntry {
    CUtensorMap tmA, tmB;
    PhaseMaps tmP;
    ConvParams p;
    int gx, gy;
};
struct GroupHeader {
    int n;
    int item_begin[kMaxGroup + 1];
    int stages;
    uint32_t slot_bytes;         // ring slot: max(a_stage + b_stage) over the members, multiple of 1024
    uint32_t acc_cols;           // TMEM columns per accumulator buffer: max(m_tiles * n_tile)
    uint32_t tmem_cols;          // allocation: power of two >= 2 * acc_cols
};

constexpr int kPGroupBias = 4096;     // floats: sum of the members' output channels (checked by the host)
constexpr int kPGroupThreads = 320;   // warp 0 TMA, warp 1 MMA + TMEM alloc, warps 2-5 / 6-9 two epilogue teams (alternate items)

__global__ void __launch_bounds__(kPGroupThreads, 1)
conv_umma_pgroup_kernel(const GroupEntry* __restrict__ table, const __grid_constant__ GroupHeader hdr) {
    extern __shared__ __align__(1024) uint8_t smem_raw[];
    __shared__ ConvParams s_p[kMaxGroup];
    __shared__ int s_gx[kMaxGroup];
    __shared__ __align__(16) float s_bias[kPGroupBias];           // all output channels of every member, member e at s_boff[e]
    __shared__ int s_boff[kMaxGroup];
    __shared__ __align__(8) unsigned long long s_bar[2 * kMaxPStages + 4];
    __shared__ uint32_t s_tmem;

    const uint32_t ring = (smem_u32(smem_raw) + 1023u) & ~1023u;
    const uint32_t full_bar = smem_u32(&s_bar[0]);
    const uint32_t empty_bar = full_bar + 8u * kMaxPStages;
    const uint32_t acc_full = empty_bar + 8u * kMaxPStages;       // 2 barriers
    const uint32_t acc_empty = acc_full + 16u;                    // 2 barriers
    const int warp = uniform_warp_idx(), lane = threadIdx.x & 31;
    const int n_members = hdr.n;
    const int total = hdr.item_begin[n_members];
    const int stages = hdr.stages;
    const uint32_t slot = hdr.slot_bytes;

    pdl_launch_dependents();             // the next launch of this stream may start its prologue
    for (int e = 0; e < n_members; ++e) {
        const uint32_t* src = reinterpret_cast<const uint32_t*>(&table[e].p);
        uint32_t* dst = reinterpret_cast<uint32_t*>(&s_p[e]);
        for (int i = threadIdx.x; i < (int)(sizeof(ConvParams) / 4); i += kPGroupThreads) dst[i] = src[i];
    }
    {
        int boff = 0;
        for (int e = 0; e < n_members; ++e) {
            const int cout = table[e].p.Cout;
            const float* bsrc = table[e].p.bias;
            for (int i = threadIdx.x; i < cout; i += kPGroupThreads) s_bias[boff + i] = bsrc[i];
            if (threadIdx.x == 0) s_boff[e] = boff;
            boff += (cout + 3) & ~3;
        }
    }
    if ((int)threadIdx.x < n_members) s_gx[threadIdx.x] = table[threadIdx.x].gx;
    if (warp == 0) {
        for (int s = lane; s < stages; s += 32) {
            mbar_init(full_bar + 8u * s, 1);
            mbar_init(empty_bar + 8u * s, 1);
        }
        if (lane < 2) {
            mbar_init(acc_full + 8u * lane, 1);
            mbar_init(acc_empty + 8u * lane, 4);                  // one arrival per epilogue warp
        }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 1) tmem_alloc(smem_u32(&s_tmem), hdr.tmem_cols);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *reinterpret_cast<volatile uint32_t*>(&s_tmem);
    pdl_wait();                          // activations / residuals / outputs below belong to earlier launches

    // work item w -> (member e, tile, output-channel split)
    auto member_of = [&](int w) {
        int e = 0;
        while (e + 1 < n_members && w >= hdr.item_begin[e + 1]) ++e;
        return e;
    };

    if (warp == 0) {
        // ===== TMA producer (warp-wide control flow, one elected lane issues) =====
        int s = 0;
        uint32_t ph = 0;
        for (int w = (int)blockIdx.x; w < total; w += (int)gridDim.x) {
            const int e = member_of(w);
            const ConvParams& p = s_p[e];
            const CUtensorMap* tmA = &table[e].tmA;
            const CUtensorMap* tmB = &table[e].tmB;
            const int local = w - hdr.item_begin[e], gx = s_gx[e];
            int t = local % gx;
            const int by = local / gx;
            const int tile_w = t % p.tiles_w; t /= p.tiles_w;
            const int tile_h = t % p.tiles_h; t /= p.tiles_h;
            const int n0 = t * p.tn, h0 = tile_h * p.th, w0 = tile_w * p.tw;
            const int n_off = by * p.n_tile;
            const int ksz = p.ksz, pad = ksz / 2, taps = ksz * ksz, n_chunks = p.n_chunks, chunk = p.chunk, stride = p.stride, Cout = p.Cout;
            const uint32_t tx = p.tx_bytes, a_bytes = p.a_stage_bytes;
            const bool phase = p.phase_maps != 0;
            if (p.tl && lane == 0) atomicMin(p.tl, globaltimer_ns());
            for (int kk = 0; kk < taps * n_chunks; ++kk) {          // K order (chunk, dx, dy): same as the halo kernel
                const int cc = kk / taps, tk = kk % taps;
                const int dxi = tk / ksz, dyi = tk % ksz, tap = dyi * ksz + dxi;
                const int dy = dyi - pad, dx = dxi - pad;
                {
                    mbar_wait(empty_bar + 8u * s, ph ^ 1u);
                    if (elect_one()) {
                        mbar_expect_tx(full_bar + 8u * s, tx);
                        if (phase)
                            tma_load_4d(ring + s * slot, &table[e].tmP.m[(dy & 1) * 2 + (dx & 1)], full_bar + 8u * s, cc * chunk,
                                        w0 - (dx < 0), h0 - (dy < 0), n0);
                        else
                            tma_load_4d(ring + s * slot, tmA, full_bar + 8u * s, cc * chunk, w0 * stride + dx, h0 * stride + dy, n0);
                        tma_load_2d(ring + s * slot + a_bytes, tmB, full_bar + 8u * s, cc * chunk, tap * Cout + n_off);
                    }
                    __syncwarp();
                    if (++s == stages) { s = 0; ph ^= 1u; }
                }
            }
        }
    } else if (warp == 1) {
        // ===== MMA issuer (warp-wide control flow, one elected lane issues) =====
        int s = 0, cnt = 0;
        uint32_t ph = 0;
        for (int w = (int)blockIdx.x; w < total; w += (int)gridDim.x, ++cnt) {
            const int e = member_of(w);
            const ConvParams& p = s_p[e];
            const int ab = cnt & 1, use = cnt >> 1;
            if (use > 0) {                                   // the epilogue of item cnt-2 has drained this buffer
                mbar_wait(acc_empty + 8u * ab, (uint32_t)((use - 1) & 1));
                tc_fence_after();
            }
            const uint32_t d_base = tmem_base + (uint32_t)ab * hdr.acc_cols;
            const int ksteps = p.chunk / 16, m_tiles = p.m_tiles, n_tile = p.n_tile;
            const int k_iters = p.ksz * p.ksz * p.n_chunks;
            const uint32_t idesc = p.idesc, a_bytes = p.a_stage_bytes;
            const uint64_t d0 = make_desc(0, p.row_bytes);
            const uint32_t mt_step = (128u * p.row_bytes) >> 4;
            for (int it = 0; it < k_iters; ++it) {
                mbar_wait(full_bar + 8u * s, ph);
                tc_fence_after();
                const uint64_t ad0 = d0 + ((ring + s * slot) >> 4);
                const uint64_t bd0 = d0 + ((ring + s * slot + a_bytes) >> 4);
                if (elect_one()) {
                    for (int mt = 0; mt < m_tiles; ++mt) {
                        const uint64_t ad = ad0 + mt * mt_step;
                        const uint32_t dt = d_base + mt * n_tile;
                        umma_f16(dt, ad, bd0, idesc, it ? 1u : 0u);
                        umma_f16(dt, ad + 2, bd0 + 2, idesc, 1u);
                        if (ksteps == 4) {
                            umma_f16(dt, ad + 4, bd0 + 4, idesc, 1u);
                            umma_f16(dt, ad + 6, bd0 + 6, idesc, 1u);
                        }
                    }
                    umma_commit(empty_bar + 8u * s);          // frees the stage when these MMAs retire
                }
                __syncwarp();
                if (++s == stages) { s = 0; ph ^= 1u; }
            }
            if (elect_one()) umma_commit(acc_full + 8u * ab);
            __syncwarp();
        }
    } else {
        // ===== epilogue: TMEM -> registers -> (+bias, +residual, ReLU) -> global =====
        // team t drains items t, t+2, ... of this CTA from accumulator buffer t, so that the global-memory
        // latency of one item's residual loads / stores overlaps the other team's item
        const int grp = warp & 3;                          // TMEM lane group this warp may read
        const int team = (warp - 2) >> 2;
        const int et = ((int)threadIdx.x - 64) & 127;      // 0..127 inside the team
        int cnt = team;
        for (int w = (int)blockIdx.x + team * (int)gridDim.x; w < total; w += 2 * (int)gridDim.x, cnt += 2) {
            const int e = member_of(w);
            const ConvParams& p = s_p[e];
            const int local = w - hdr.item_begin[e], gx = s_gx[e];
            int t = local % gx;
            const int by = local / gx;
            const int tile_w = t % p.tiles_w; t /= p.tiles_w;
            const int tile_h = t % p.tiles_h; t /= p.tiles_h;
            const int n0 = t * p.tn, h0 = tile_h * p.th, w0 = tile_w * p.tw;
            const int n_tile = p.n_tile, n_off = by * n_tile, m_tiles = p.m_tiles, up = p.up, Cout = p.Cout, relu = p.relu;
            const __half* res = p.res;
            __half* out = p.out;
            const int ab = cnt & 1, use = cnt >> 1;
            const float* bias_s = s_bias + s_boff[e] + n_off;
            const int Hout = p.Ho * up, Wout = p.Wo * up;
            bool valid[2];
            int pn[2], ph_[2], pw[2];
#pragma unroll
            for (int mt = 0; mt < 2; ++mt) {
                const int R = mt * 128 + grp * 32 + lane;
                pw[mt] = w0 + R % p.tw; ph_[mt] = h0 + (R / p.tw) % p.th; pn[mt] = n0 + R / (p.tw * p.th);
                valid[mt] = mt < m_tiles && pn[mt] < p.P;
            }
            const bool direct = up == 1;
            uint4 rq[8];
            auto fetch_res = [&](int mt, int cg) {
                const size_t o = ((((size_t)pn[mt] * Hout + ph_[mt]) * Wout) + pw[mt]) * Cout + n_off + cg;
#pragma unroll
                for (int q = 0; q < 8; ++q)
                    if (cg + q * 8 < n_tile) rq[q] = *reinterpret_cast<const uint4*>(res + o + q * 8);
            };
            const int dbg = p.dbg_flags;                       // bring-up (HBP_PG_DBG): 1 no stores, 2 no residual, 8 no epilogue work at all
            if (dbg & 2) res = nullptr;
            if (direct && res && valid[0]) fetch_res(0, 0);
            mbar_wait(acc_full + 8u * ab, (uint32_t)(use & 1));
            tc_fence_after();
            const uint32_t t_base = tmem_base + ((uint32_t)(grp * 32) << 16) + (uint32_t)ab * hdr.acc_cols;
            for (int mt = 0; mt < ((dbg & 8) ? 0 : m_tiles); ++mt) {
                for (int cg = 0; cg < n_tile; cg += 64) {
                    if (direct && res && valid[mt] && (mt | cg)) fetch_res(mt, cg);
#pragma unroll
                    for (int cq = 0; cq < 4; ++cq) {
                        const int c0 = cg + cq * 16;
                        if (c0 >= n_tile) break;
                        uint32_t r[16];
                        tmem_ld16(t_base + (uint32_t)(mt * n_tile + c0), r);
                        tmem_ld_wait();
                        if (!valid[mt]) continue;
                        float v[16];
#pragma unroll
                        for (int q = 0; q < 16; ++q) v[q] = __uint_as_float(r[q]) + bias_s[c0 + q];
                        for (int uy = 0; uy < up; ++uy)
                            for (int ux = 0; ux < up; ++ux) {
                                const size_t o = ((((size_t)pn[mt] * Hout + ph_[mt] * up + uy) * Wout) + pw[mt] * up + ux) * Cout + n_off + c0;
                                float x[16];
#pragma unroll
                                for (int q = 0; q < 16; ++q) x[q] = v[q];
                                if (res) {
                                    uint4 q0, q1;
                                    if (direct) { q0 = rq[2 * cq]; q1 = rq[2 * cq + 1]; }
                                    else {
                                        q0 = *reinterpret_cast<const uint4*>(res + o);
                                        q1 = *reinterpret_cast<const uint4*>(res + o + 8);
                                    }
                                    const __half2* h0p = reinterpret_cast<const __half2*>(&q0);
                                    const __half2* h1p = reinterpret_cast<const __half2*>(&q1);
#pragma unroll
                                    for (int q = 0; q < 4; ++q) {
                                        const float2 f0 = __half22float2(h0p[q]), f1 = __half22float2(h1p[q]);
                                        x[2 * q] += f0.x; x[2 * q + 1] += f0.y;
                                        x[8 + 2 * q] += f1.x; x[8 + 2 * q + 1] += f1.y;
                                    }
                                }
                                if (relu) {
#pragma unroll
                                    for (int q = 0; q < 16; ++q) x[q] = fmaxf(x[q], 0.f);
                                }
                                __align__(16) __half2 pk[8];
#pragma unroll
                                for (int q = 0; q < 8; ++q) pk[q] = __floats2half2_rn(x[2 * q], x[2 * q + 1]);
                                if ((dbg & 1) && pk[0].x != __float2half(12345.f)) continue;
                                store_half16(out + o, pk);
                            }
                    }
                }
            }
            // accumulator buffer back to the MMA warp
            tc_fence_before();
            __syncwarp();
            if (lane == 0) asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(acc_empty + 8u * ab) : "memory");
            if (p.tl && et == 0) atomicMax(p.tl + 1, globaltimer_ns());
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 1) {
        tc_fence_after();
        tmem_dealloc(tmem_base, hdr.tmem_cols);
    }
}

constexpr int kHaloW = 10;       // 8 output columns + 1 halo column each side
constexpr int kMaxChunks = 8;
constexpr int kMaxAStages = 4;   // halo tiles in flight per CTA
constexpr int kMaxBSlots = 24;   // (3 dx) x 8 Cin chunks: weight slots of three taps each (resident) or ring depth (streamed)
constexpr int kHaloThreads = 384;   // warp 0 halo + weight TMA, warps 1-2 MMA issuers (warp 1 owns TMEM), warp 3 residual TMA, warps 4-7 / 8-11 two epilogue teams
constexpr int kMaxAccBufs = 4;
constexpr uint32_t kHaloBarBytes = 8u * (2 * kMaxAStages * kMaxChunks + 3 * kMaxAStages + 2 * kMaxBSlots + 2 * kMaxAccBufs) + 16u;

// 3x3 stride-1 convolution, halo mode, one persistent CTA per SM.  Tile = tn images x th rows
// x 8 columns.  In shared memory a chunk is the TMA box (chunk channels, 10, rs = th+2, tn):
// pixel rows of `row_bytes` ordered [n][h][w], i.e. "stacked" image rows q = n*rs + h of 10
// pixels each.  MMA row r of M-tile mt is pixel column r%8 of stacked row g = mt*16 + r/8; tap
// (dy,dx) reads stacked row g+dy, column r%8+dx: a K-major operand with start offset
// ((mt*16+dy)*10+dx)*row_bytes and an 8-row-group stride of 10 rows.  Stacked rows that are
// halo rows (g % rs >= th) yield garbage accumulator rows which the epilogue skips.
//
// Pipeline (all phases of consecutive tiles overlap inside ONE CTA -- co-resident CTAs of one
// launch run in lock step and do not hide each other's phases):
//   * the halo tiles go through a ring of p.a_stages buffers: the producer warp runs up to
//     a_stages tiles ahead of the MMA lane;
//   * the weights of all nine taps stay resident in shared memory for the CTA's lifetime when
//     they fit (p.b_resident), otherwise they stream through a deep ring of p.b_slots stages;
//   * the accumulator is double-buffered in TMEM (p.acc_bufs == 2) and two teams of four
//     epilogue warps drain alternate tiles while the MMA lane fills the other buffer;
//   * the tensor pipe queues only ~2 MMAs, so every cycle the issuing lane spends on barrier
//     waits, fences and commits between tiles idles the pipe: with resident weights TWO warps
//     issue, tile j by warp j % issuers into accumulator buffer j % acc_bufs, and hide each
//     other's per-tile overhead;
//   * one warp saturates the tensor pipe with M=128 x N=32 MMAs (40 cycles each, shared-memory
//     feed bound; tools/umma_rate_probe.cu) only if nothing but the MMAs sits in its issue
//     path: the kernel is a template on (k-steps per chunk, M-tiles) and the MMAs of one weight
//     slot (three taps) are straight-line code whose descriptors differ by immediates.
// Each CTA walks tiles blockIdx.x, +gridDim.x, ... (grid <= SM count: no partial wave).
constexpr int kResPrefetch = 8;   // uint4 (8 halfs) of residual a thread may hold in flight

template <int KSTEPS, int MT, int KS>
__global__ void __launch_bounds__(kHaloThreads, 1)
conv_umma_halo_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
                      const __grid_constant__ CUtensorMap tmR, const ConvParams p) {
    constexpr int HW = KS == 3 ? kHaloW : 8;         // pixels per stacked row of the tile in shared memory (halo columns for 3x3)
    extern __shared__ __align__(1024) uint8_t smem_raw[];
    if (p.dbg_flags & 4) return;         // bring-up: empty launch (measures the launch / dependency floor of the graph)
    const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
    const uint32_t a_tile_bytes = p.n_chunks * p.a_chunk_bytes;
    const uint32_t a_base = smem_base;
    const uint32_t b_base = a_base + p.a_stages * a_tile_bytes;
    const uint32_t r_base = b_base + p.b_slots * (uint32_t)KS * p.b_stage_bytes;  // residual ring: a_stages x r_chunks x r_chunk_bytes
    const uint32_t r_tile_bytes = p.res_smem ? p.r_chunks * p.r_chunk_bytes : 0u;
    // constant operand tiles, 64-byte rows in the SWIZZLE_64B pattern (1024-aligned): ones (128 x K32: K columns 0,1 = 1),
    // identity (32 x K32), bias (n_tile x K32: K column 0 = fp16(bias), 1 = fp16(bias - column 0))
    const uint32_t c_base = r_base + p.a_stages * r_tile_bytes;
    const uint32_t c_ones = c_base, c_ident = c_base + 8192u, c_bias = c_base + 10240u;
    const uint32_t bar_base = c_base + p.c_bytes;
    const uint32_t a_full = bar_base;                                   // a_stages x n_chunks
    const uint32_t a_empty = a_full + 8u * (kMaxAStages * kMaxChunks);  // a_stages
    const uint32_t b_full = a_empty + 8u * kMaxAStages;                 // b_slots
    const uint32_t b_empty = b_full + 8u * kMaxBSlots;                  // b_slots (ring only)
    const uint32_t acc_full = b_empty + 8u * kMaxBSlots;                // acc_bufs
    const uint32_t acc_empty = acc_full + 8u * kMaxAccBufs;             // acc_bufs
    const uint32_t res_full = acc_empty + 8u * kMaxAccBufs;             // a_stages
    const uint32_t res_empty = res_full + 8u * kMaxAStages;             // a_stages
    const uint32_t a_cempty = res_empty + 8u * kMaxAStages;             // a_stages x n_chunks (streamed weights: chunks are handed back one by one)
    const uint32_t tmem_slot = a_cempty + 8u * (kMaxAStages * kMaxChunks);
    float* s_bias = reinterpret_cast<float*>(smem_raw + (bar_base + kHaloBarBytes - smem_u32(smem_raw)));

    const int warp = uniform_warp_idx(), lane = threadIdx.x & 31;
    const int n_off = blockIdx.y * p.n_tile;
    const int tiles_per_img = p.tiles_w * p.tiles_h;
    long long* dbg = (p.dbg && blockIdx.x < 64 && blockIdx.y == 0) ? p.dbg + blockIdx.x * 256 : nullptr;
    if (dbg && threadIdx.x == 0) dbg[0] = clock64();
    pdl_launch_dependents();             // the next launch of this stream may start its prologue
    if (p.tl && threadIdx.x == 0) atomicMin(p.tl, globaltimer_ns());

    if (warp == 0) {
        // barrier init spread over the lanes of warp 0 (one thread doing ~60 inits is a microsecond of prologue)
        if (lane == 0) {
            asm volatile("prefetch.tensormap [%0];" ::"l"(&tmA) : "memory");
            asm volatile("prefetch.tensormap [%0];" ::"l"(&tmB) : "memory");
            if (p.res_smem) asm volatile("prefetch.tensormap [%0];" ::"l"(&tmR) : "memory");
        }
        for (int i = lane; i < p.a_stages * p.n_chunks; i += 32) { mbar_init(a_full + 8u * i, 1); mbar_init(a_cempty + 8u * i, 1); }
        for (int i = lane; i < p.a_stages; i += 32) { mbar_init(a_empty + 8u * i, 1); mbar_init(res_full + 8u * i, 1); mbar_init(res_empty + 8u * i, p.res_mma ? 1 : 4); }
        for (int i = lane; i < p.b_slots; i += 32) { mbar_init(b_full + 8u * i, 1); mbar_init(b_empty + 8u * i, 1); }
        for (int i = lane; i < kMaxAccBufs; i += 32) { mbar_init(acc_full + 8u * i, 1); mbar_init(acc_empty + 8u * i, 4); }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 1) tmem_alloc(tmem_slot, p.tmem_cols);
    for (int i = threadIdx.x; i < p.n_tile; i += kHaloThreads) s_bias[i] = p.bias[n_off + i];
    if (p.c_bytes) {
        // element (row r, K column k) of a 64-byte-row SWIZZLE_64B tile: 16-byte chunk (k >> 3) ^ ((r >> 1) & 3) of row r
        for (uint32_t i = threadIdx.x; i < p.c_bytes / 16u; i += kHaloThreads)
            asm volatile("st.shared.v4.b32 [%0], {%1, %1, %1, %1};" ::"r"(c_base + 16u * i), "r"(0u) : "memory");
        __syncthreads();
        const uint32_t one2 = 0x3C003C00u;                                  // (1.0h, 1.0h)
        for (uint32_t r = threadIdx.x; r < 128u; r += kHaloThreads)
            asm volatile("st.shared.b32 [%0], %1;" ::"r"(c_ones + r * 64u + (((r >> 1) & 3u) << 4)), "r"(one2) : "memory");
        for (uint32_t n = threadIdx.x; n < 32u; n += kHaloThreads)
            asm volatile("st.shared.b16 [%0], %1;" ::"r"(c_ident + n * 64u + ((((n >> 3) ^ (n >> 1)) & 3u) << 4) + (n & 7u) * 2u), "h"((unsigned short)0x3C00) : "memory");
        for (uint32_t n = threadIdx.x; n < (uint32_t)p.n_tile; n += kHaloThreads) {
            const float b = p.bias[n_off + n];
            const __half hi = __float2half_rn(b);
            const __half lo = __float2half_rn(b - __half2float(hi));
            const uint32_t pk = (uint32_t)__half_as_ushort(hi) | ((uint32_t)__half_as_ushort(lo) << 16);
            asm volatile("st.shared.b32 [%0], %1;" ::"r"(c_bias + n * 64u + (((n >> 1) & 3u) << 4)), "r"(pk) : "memory");
        }
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");       // generic-proxy writes -> tcgen05.mma (async proxy) reads
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    uint32_t tmem_base;
    asm volatile("ld.shared.b32 %0, [%1];" : "=r"(tmem_base) : "r"(tmem_slot));
    tmem_base = __shfl_sync(0xffffffffu, tmem_base, 0);
    if (dbg && threadIdx.x == 0) dbg[1] = clock64();
    const uint32_t acc_stride = (uint32_t)(MT * p.n_tile);              // TMEM columns per accumulator buffer
    const uint32_t b_bytes = (uint32_t)p.n_tile * p.row_bytes;          // one tap; a slot holds three (dy = 0..2 of one dx)
    const uint32_t b_slot_bytes = (uint32_t)KS * p.b_stage_bytes;

    if (warp == 0) {
        // ===== TMA producer (warp-wide control flow, one elected lane issues) =====
        const int T = (p.n_tiles - (int)blockIdx.x + (int)gridDim.x - 1) / (int)gridDim.x;   // tiles of this CTA
        // halo tile t of this CTA -> ring stage t % a_stages; the stage is free once the MMAs of
        // tile t - a_stages have retired (completion number t / a_stages - 1 of its a_empty barrier)
        TileWalk aw;                             // coordinates of the next tile whose halo is requested
        aw.init((int)blockIdx.x, (int)gridDim.x, p.tiles_w, p.tiles_h, p.n_tiles, p.reverse);
        int a_sa = 0, a_lap = 0;
        auto issue_a = [&](int t) {             // called for t = 0, 1, 2, ... in order
            (void)t;
            const int sa = a_sa, lap = a_lap;
            const int n0 = aw.tg * p.tn, h0 = aw.th * p.th, w0 = aw.tw * 8;
            if (lap > 0) mbar_wait(a_empty + 8u * sa, (uint32_t)((lap - 1) & 1));
            if (dbg && lane == 0 && t < 16) dbg[128 + t] = clock64();     // A(t) requested
            if (elect_one()) {
                for (int cc = 0; cc < p.n_chunks; ++cc) {
                    const uint32_t bar = a_full + 8u * (sa * p.n_chunks + cc);
                    mbar_expect_tx(bar, p.a_box_bytes);
                    tma_load_4d_hint(a_base + sa * a_tile_bytes + cc * p.a_chunk_bytes, &tmA, bar, cc * p.chunk, w0 - KS / 2, h0 - KS / 2, n0, p.hint_a);
                }
            }
            __syncwarp();

            if (++a_sa == p.a_stages) { a_sa = 0; ++a_lap; }
            aw.next();
        };
        // weights do not depend on the previous launch: request them before the dependency wait
        if (p.b_resident) {
            // all weights of this CTA's output-channel slice, once; one barrier per (chunk, tap)
            // so that the first tile starts as soon as its first tap has landed
            if (elect_one()) {
                for (int cc = 0; cc < p.n_chunks; ++cc)
                    for (int dx = 0; dx < KS; ++dx) {
                        const int i = cc * KS + dx;
                        mbar_expect_tx(b_full + 8u * i, (uint32_t)KS * b_bytes);
                        for (int dy = 0; dy < KS; ++dy)
                            tma_load_2d(b_base + i * b_slot_bytes + dy * p.b_stage_bytes, &tmB, b_full + 8u * i, cc * p.chunk,
                                        (dy * KS + dx) * p.Cout + n_off);
                    }
            }
            __syncwarp();
        }
        // streamed weights (they do not fit beside a halo tile: the 128-channel branch, transition1): this warp does
        // nothing but keep the weight ring full -- the halo tiles come from warp 3, chunk by chunk.  (With both on one warp
        // the wait for a free halo stage sat between the weight requests of consecutive tiles: every tile started with an
        // empty ring.)  Off by default (HBP_HALO_SPLIT_PRODUCER=1): measured gains only with the conv alone on the GPU.
        const bool split = !p.b_resident && p.split_producer;
        if (!split) {
            pdl_wait();                      // activations below are the previous launch's output
            for (int t = 0; t < p.a_stages && t < T; ++t) issue_a(t);
        }
        int s = 0;
        uint32_t ph = 0;
        for (int j = 0; j < T; ++j) {
            if (!p.b_resident) {
                for (int cc = 0; cc < p.n_chunks; ++cc)
                    for (int dx = 0; dx < KS; ++dx) {
                        mbar_wait(b_empty + 8u * s, ph ^ 1u);
                        if (elect_one()) {
                            mbar_expect_tx(b_full + 8u * s, (uint32_t)KS * b_bytes);
                            for (int dy = 0; dy < KS; ++dy)
                                tma_load_2d(b_base + s * b_slot_bytes + dy * p.b_stage_bytes, &tmB, b_full + 8u * s, cc * p.chunk,
                                            (dy * KS + dx) * p.Cout + n_off);
                        }
                        __syncwarp();
                        if (++s == p.b_slots) { s = 0; ph ^= 1u; }
                    }
            }
            if (!split && j + p.a_stages < T) issue_a(j + p.a_stages);
        }
    } else if (warp == 3) {
        if (!p.b_resident && p.split_producer) {
            // ===== halo + residual producer of the streamed-weights mode: chunk cc of stage sa is refilled as soon as
            // the MMAs that read it have retired (a_cempty), i.e. a part of a tile ahead even with a single stage =====
            const int T = (p.n_tiles - (int)blockIdx.x + (int)gridDim.x - 1) / (int)gridDim.x;
            TileWalk aw;
            aw.init((int)blockIdx.x, (int)gridDim.x, p.tiles_w, p.tiles_h, p.n_tiles, p.reverse);
            int sa = 0, lap = 0, sr = 0, lap_r = 0;
            pdl_wait();
            for (int t = 0; t < T; ++t) {
                const int n0 = aw.tg * p.tn, h0 = aw.th * p.th, w0 = aw.tw * 8;
                for (int cc = 0; cc < p.n_chunks; ++cc) {
                    const int i = sa * p.n_chunks + cc;
                    if (lap > 0) mbar_wait(a_cempty + 8u * i, (uint32_t)((lap - 1) & 1));
                    if (elect_one()) {
                        mbar_expect_tx(a_full + 8u * i, p.a_box_bytes);
                        tma_load_4d_hint(a_base + sa * a_tile_bytes + cc * p.a_chunk_bytes, &tmA, a_full + 8u * i, cc * p.chunk, w0 - KS / 2, h0 - KS / 2, n0, p.hint_a);
                    }
                    __syncwarp();
                    if (cc == 0 && p.res_smem) {
                        // the tile's residual rows right behind its first chunk
                        if (lap_r > 0) mbar_wait(res_empty + 8u * sr, (uint32_t)((lap_r - 1) & 1));
                        if (elect_one()) {
                            mbar_expect_tx(res_full + 8u * sr, (uint32_t)p.r_chunks * p.r_box_bytes);
                            for (int rc = 0; rc < p.r_chunks; ++rc)
                                tma_load_4d_hint(r_base + sr * r_tile_bytes + rc * p.r_chunk_bytes, &tmR, res_full + 8u * sr, n_off + rc * 64,
                                                 aw.tw * 8, aw.th * p.th, aw.tg * p.tn, p.hint_r);
                        }
                        __syncwarp();
                        if (++sr == p.a_stages) { sr = 0; ++lap_r; }
                    }
                }
                if (++sa == p.a_stages) { sa = 0; ++lap; }
                aw.next();
            }
        } else
        // ===== residual producer: the residual rows of every tile (no halo), 64 output channels per box,
        // through their own ring so that a slow epilogue never delays the halo requests =====
        if (p.res_smem) {
            const int T = (p.n_tiles - (int)blockIdx.x + (int)gridDim.x - 1) / (int)gridDim.x;
            TileWalk rw;
            rw.init((int)blockIdx.x, (int)gridDim.x, p.tiles_w, p.tiles_h, p.n_tiles, p.reverse);
            int sr = 0, lap = 0;
            pdl_wait();
            for (int t = 0; t < T; ++t) {
                if (lap > 0) mbar_wait(res_empty + 8u * sr, (uint32_t)((lap - 1) & 1));
                if (elect_one()) {
                    mbar_expect_tx(res_full + 8u * sr, (uint32_t)p.r_chunks * p.r_box_bytes);
                    for (int rc = 0; rc < p.r_chunks; ++rc)
                        tma_load_4d_hint(r_base + sr * r_tile_bytes + rc * p.r_chunk_bytes, &tmR, res_full + 8u * sr, n_off + rc * 64,
                                         rw.tw * 8, rw.th * p.th, rw.tg * p.tn, p.hint_r);
                }
                __syncwarp();
                if (++sr == p.a_stages) { sr = 0; ++lap; }
                rw.next();
            }
        }
    } else if (warp <= 2) {
      if (warp - 1 < p.issuers) {
        // ===== MMA issuers (warp-wide control flow, one elected lane issues) =====
        const int iss = warp - 1;
        constexpr uint32_t kRow16 = KSTEPS * 2;                                 // one pixel row (chunk halfs) in 16-byte units
        constexpr uint32_t kMtStep = 16u * HW * kRow16;
        const uint64_t da0 = make_desc(0, p.row_bytes, HW * p.row_bytes);   // A: 8-row groups one stacked row (HW pixels) apart
        const uint64_t db0 = make_desc(0, p.row_bytes);
        const uint64_t dc0 = make_desc(0, 64);                                  // constant tiles: 64-byte rows, SWIZZLE_64B
        const uint64_t dr0 = make_desc(0, p.res_smem ? p.r_row_bytes : 64u);    // residual tile as an A operand
        const uint32_t bstep = p.b_stage_bytes >> 4;                            // one tap of weights in 16-byte units
        const uint32_t idesc = p.idesc, n_tile = (uint32_t)p.n_tile;
        const int T = (p.n_tiles - (int)blockIdx.x + (int)gridDim.x - 1) / (int)gridDim.x;   // tiles of this CTA
        int s = 0;
        uint32_t ph = 0;
        int sa = iss % p.a_stages, lap_a = iss / p.a_stages;      // halo stage / lap of tile j
        int ab = iss % p.acc_bufs, use = iss / p.acc_bufs;        // accumulator buffer of tile j / how often it was filled before
        for (int j = iss; j < T; j += p.issuers) {
            const uint32_t pa = (uint32_t)(lap_a & 1);
            // accumulator buffer `ab` must have been drained by the epilogue of its previous tile
            if (use > 0) {
                mbar_wait(acc_empty + 8u * ab, (uint32_t)((use - 1) & 1));
                tc_fence_after();
            }
            const uint32_t d_base = tmem_base + ab * acc_stride;
            if (dbg && lane == 0 && j < 16) dbg[32 + 3 * j] = clock64();        // accumulator free
            if (p.bias_mma) {
                // D = ones x [bias_hi | bias_lo]^T: overwrites the buffer, needs neither the halo tile nor the weights
                if (!(p.dbg_flags & 16) && elect_one()) {
#pragma unroll
                    for (int mt = 0; mt < MT; ++mt) umma_f16(d_base + mt * n_tile, dc0 + (c_ones >> 4), dc0 + (c_bias >> 4), idesc, 0u);
                }
                __syncwarp();
            }
            for (int cc = 0; cc < p.n_chunks; ++cc) {
                mbar_wait(a_full + 8u * (sa * p.n_chunks + cc), pa);
                if (dbg && cc == 0 && j == 0 && lane == 0) dbg[2] = clock64();
                if (dbg && cc == 0 && lane == 0 && j < 16) dbg[33 + 3 * j] = clock64();   // A landed
                const uint64_t a_c = da0 + ((a_base + sa * a_tile_bytes + cc * p.a_chunk_bytes) >> 4);
                if (p.b_resident && j >= p.issuers) {
                    // steady state with resident weights: nothing to wait for between the taps -- the
                    // 9 x KSTEPS x MT MMAs of the chunk are one straight-line block under one election
                    // (every instruction between two MMAs idles the pipe: it queues only ~2 of them)
                    tc_fence_after();
                    const uint64_t b_c = db0 + ((b_base + (uint32_t)(cc * KS) * b_slot_bytes) >> 4);
                    const uint32_t fresh = (cc == 0 && !p.bias_mma) ? 0u : 1u;
                    if (!(p.dbg_flags & 16) && elect_one()) {
#pragma unroll
                        for (int dx = 0; dx < KS; ++dx)
#pragma unroll
                            for (int dy = 0; dy < KS; ++dy)
#pragma unroll
                                for (int ks = 0; ks < KSTEPS; ++ks)
#pragma unroll
                                    for (int mt = 0; mt < MT; ++mt)
                                        umma_f16(d_base + mt * n_tile, a_c + (uint32_t)((dy * HW + dx) * kRow16 + mt * kMtStep + 2 * ks),
                                                 b_c + (uint32_t)(dx * KS + dy) * bstep + 2 * ks, idesc, (dx | dy | ks) ? 1u : fresh);
                    }
                    __syncwarp();
                    continue;
                }
#pragma unroll
                for (int dx = 0; dx < KS; ++dx) {
                    const int slot = p.b_resident ? cc * KS + dx : s;
                    if (!p.b_resident) mbar_wait(b_full + 8u * s, ph);
                    else if (j < p.issuers) mbar_wait(b_full + 8u * slot, 0);      // this warp's first tile
                    if (dbg && cc == 0 && dx == 0 && j == 0 && lane == 0) dbg[3] = clock64();
                    tc_fence_after();
                    const uint64_t b_s = db0 + ((b_base + slot * b_slot_bytes) >> 4);
                    const uint32_t fresh = (cc == 0 && dx == 0 && !p.bias_mma) ? 0u : 1u;      // first MMA of a tile overwrites
                    if (elect_one()) {
#pragma unroll
                        for (int dy = 0; dy < KS; ++dy) {
                            const uint64_t bd = b_s + (uint32_t)dy * bstep;
#pragma unroll
                            for (int ks = 0; ks < KSTEPS; ++ks)
#pragma unroll
                                for (int mt = 0; mt < MT; ++mt)
                                    umma_f16(d_base + mt * n_tile, a_c + (uint32_t)((dy * HW + dx) * kRow16 + mt * kMtStep + 2 * ks),
                                             bd + 2 * ks, idesc, (dy | ks) ? 1u : fresh);
                        }
                        if (!p.b_resident) umma_commit(b_empty + 8u * s);
                    }
                    __syncwarp();
                    if (!p.b_resident && ++s == p.b_slots) { s = 0; ph ^= 1u; }
                }
                if (!p.b_resident && p.split_producer) {
                    if (elect_one()) umma_commit(a_cempty + 8u * (sa * p.n_chunks + cc));     // this chunk of the halo stage is free when its MMAs retire
                    __syncwarp();
                }
            }
            if (p.res_mma) {
                // D[:, 32j .. 32j+32) += residual[:, 32j .. 32j+32) x I32: the residual tile of the TMA ring (pixel rows in
                // the hardware swizzle, the A-operand layout) against the identity; exact (fp16 x 1.0 into fp32)
                mbar_wait(res_full + 8u * sa, pa);
                tc_fence_after();
                if (elect_one()) {
                    const uint32_t rst = r_base + sa * r_tile_bytes;
                    const uint32_t r_ch = p.r_row_bytes >> 1;                     // channels per residual box (32 or 64)
                    for (int mt = 0; mt < ((p.dbg_flags & 16) ? 0 : MT); ++mt)
                        for (uint32_t jb = 0; jb < n_tile / 32u; ++jb) {
                            const uint32_t ch = jb * 32u;
                            const uint32_t box = rst + (ch / r_ch) * p.r_chunk_bytes + (uint32_t)mt * 128u * p.r_row_bytes;
                            const uint64_t ad = dr0 + ((box + (ch % r_ch) * 2u) >> 4);
                            const uint32_t dt = d_base + mt * n_tile + ch;
                            umma_f16(dt, ad, dc0 + (c_ident >> 4), p.idesc32, 1u);
                            umma_f16(dt, ad + 2, dc0 + (c_ident >> 4) + 2, p.idesc32, 1u);
                        }
                    umma_commit(res_empty + 8u * sa);    // residual stage free when these MMAs retire
                }
                __syncwarp();
            }
            if (elect_one()) {
                umma_commit(a_empty + 8u * sa);          // halo stage free when these MMAs retire
                umma_commit(acc_full + 8u * ab);         // accumulator ready for its epilogue team
            }
            __syncwarp();
            if (dbg && j == 0 && lane == 0) dbg[4] = clock64();
            if (dbg && lane == 0 && j < 16) dbg[34 + 3 * j] = clock64();        // MMAs issued
            sa += p.issuers; while (sa >= p.a_stages) { sa -= p.a_stages; ++lap_a; }
            ab += p.issuers; while (ab >= p.acc_bufs) { ab -= p.acc_bufs; ++use; }
        }
      }
    } else {
        // ===== epilogue teams: TMEM -> registers -> (+bias, +residual, ReLU) -> global =====
        // Team t drains tiles t, t + teams, ... of this CTA.  An accumulator of <= 64 columns is
        // copied to registers in one go and handed back to the MMA warps at once; the residual
        // rows come from the TMA ring in shared memory (the producer requested them together with
        // the tile's halo, several tiles ahead).  No integer division on the per-tile path: the row geometry of a thread is
        // loop-invariant and the tile coordinates advance by a precomputed (dw, dh, dn) step.
        const int team = (warp - 4) >> 2;
        const int grp = warp & 3;                               // TMEM lane group this warp may read
        const int r = grp * 32 + lane;
        const bool has_res = p.res != nullptr && !p.res_mma;      // (res_mma: the MMA warps already added it)
        const bool res_smem = p.res_smem != 0 && !p.res_mma;
        const bool add_bias = !p.bias_mma;
        const uint32_t t_lane0 = tmem_base + ((uint32_t)(grp * 32) << 16);
        const int teams = p.teams;
        const int T = team < teams ? (p.n_tiles - (int)blockIdx.x + (int)gridDim.x - 1) / (int)gridDim.x : 0;
        // loop-invariant geometry of this thread's (up to two) accumulator rows
        bool row_ok[2];
        int row_n[2], row_w;
        size_t row_off[2];
        uint32_t rs_off[2], rs_xor[2];                 // residual ring: byte offset of this thread's row in a box, its swizzle term
        row_w = r & 7;
#pragma unroll
        for (int mt = 0; mt < 2; ++mt) {
            const int g = mt * 16 + (r >> 3);
            const int nn = g / p.rs, hh = g - nn * p.rs;
            row_ok[mt] = mt < MT && nn < p.tn && hh < p.th;
            row_n[mt] = nn;
            row_off[mt] = (((size_t)nn * p.Ho + hh) * p.Wo + row_w) * p.Cout + n_off;
            const uint32_t rrow = row_ok[mt] ? (uint32_t)((nn * p.th + hh) * 8 + row_w) : 0u;   // box rows are ordered [n][h][w]
            rs_off[mt] = rrow * p.r_row_bytes;
            rs_xor[mt] = p.r_row_bytes == 128 ? (rrow & 7u) : ((rrow >> 1) & 3u);    // SWIZZLE_128B / SWIZZLE_64B on 1024-aligned boxes
        }
        // 16 residual halfs (two 16-byte chunks) of columns c0..c0+15 of this thread's row, from the ring stage at `rst`
        // (every per-M-tile quantity is selected with ?: -- indexing the two-element arrays with a run-time `mt` put them
        // in local memory, and the epilogue then waited on two dependent LDL round trips per 32 columns: r01c ncu source page)
        auto res_lds = [&](uint32_t rst, int mt, int c0, uint4& q0, uint4& q1) {
            const uint32_t rso = mt ? rs_off[1] : rs_off[0], rsx = mt ? rs_xor[1] : rs_xor[0];
            const uint32_t box = rst + (uint32_t)(c0 >> 6) * p.r_chunk_bytes + rso;
            const uint32_t ci = (uint32_t)(c0 & 63) >> 3;
            const uint32_t a0 = box + (((ci) ^ rsx) << 4), a1 = box + (((ci + 1) ^ rsx) << 4);
            asm volatile("ld.shared.v4.b32 {%0, %1, %2, %3}, [%4];" : "=r"(q0.x), "=r"(q0.y), "=r"(q0.z), "=r"(q0.w) : "r"(a0));
            asm volatile("ld.shared.v4.b32 {%0, %1, %2, %3}, [%4];" : "=r"(q1.x), "=r"(q1.y), "=r"(q1.z), "=r"(q1.w) : "r"(a1));
        };
        // tile walk: tile index -> (column tile, row tile, image group), advanced without divisions
        TileWalk ew;
        ew.init((int)blockIdx.x + team * (int)gridDim.x, teams * (int)gridDim.x, p.tiles_w, p.tiles_h, p.n_tiles, p.reverse);
        bool valid[2];
        size_t obase[2];
        auto locate = [&]() {
            const int n0 = ew.tg * p.tn, h0 = ew.th * p.th, w0 = ew.tw * 8;
            const size_t origin = (((size_t)n0 * p.Ho + h0) * p.Wo + w0) * p.Cout;
#pragma unroll
            for (int mt = 0; mt < 2; ++mt) {
                valid[mt] = row_ok[mt] && n0 + row_n[mt] < p.P && w0 + row_w < p.Wo;
                obase[mt] = origin + row_off[mt];
            }
        };
        auto advance = [&]() { ew.next(); };
        // x = accumulator columns c0..c0+15 -> +bias (+residual q0|q1) -> ReLU -> fp16 -> global
        auto finish16 = [&](const uint32_t* rr, size_t o, int c0, uint4 q0, uint4 q1) {
            float x[16];
#pragma unroll
            for (int q4 = 0; q4 < 4; ++q4) {
                const float4 b4 = (!add_bias || (p.dbg_flags & 2)) ? make_float4(0.f, 0.f, 0.f, 0.f) : *reinterpret_cast<const float4*>(s_bias + c0 + 4 * q4);
                x[4 * q4] = __uint_as_float(rr[4 * q4]) + b4.x; x[4 * q4 + 1] = __uint_as_float(rr[4 * q4 + 1]) + b4.y;
                x[4 * q4 + 2] = __uint_as_float(rr[4 * q4 + 2]) + b4.z; x[4 * q4 + 3] = __uint_as_float(rr[4 * q4 + 3]) + b4.w;
            }
            if (has_res) {
                const __half2* h0p = reinterpret_cast<const __half2*>(&q0);
                const __half2* h1p = reinterpret_cast<const __half2*>(&q1);
#pragma unroll
                for (int q = 0; q < 4; ++q) {
                    const float2 f0 = __half22float2(h0p[q]), f1 = __half22float2(h1p[q]);
                    x[2 * q] += f0.x; x[2 * q + 1] += f0.y;
                    x[8 + 2 * q] += f1.x; x[8 + 2 * q + 1] += f1.y;
                }
            }
            if (p.relu) {
#pragma unroll
                for (int q = 0; q < 16; ++q) x[q] = fmaxf(x[q], 0.f);
            }
            __align__(16) __half2 pk[8];
#pragma unroll
            for (int q = 0; q < 8; ++q) pk[q] = __floats2half2_rn(x[2 * q], x[2 * q + 1]);
            if ((p.dbg_flags & 1) && pk[0].x != __float2half(12345.f)) return;
            const uint4 v0 = *reinterpret_cast<const uint4*>(&pk[0]), v1 = *reinterpret_cast<const uint4*>(&pk[4]);
            if (p.hint_o) {
                asm volatile("st.global.L2::cache_hint.v4.b32 [%0], {%1, %2, %3, %4}, %5;" ::"l"(p.out + o), "r"(v0.x), "r"(v0.y), "r"(v0.z), "r"(v0.w), "l"(p.hint_o) : "memory");
                asm volatile("st.global.L2::cache_hint.v4.b32 [%0], {%1, %2, %3, %4}, %5;" ::"l"(p.out + o + 8), "r"(v1.x), "r"(v1.y), "r"(v1.z), "r"(v1.w), "l"(p.hint_o) : "memory");
            } else if (p.dbg_flags & 32) {           // bring-up: two 16-byte stores instead of one 32-byte store
                *reinterpret_cast<uint4*>(p.out + o) = v0;
                *reinterpret_cast<uint4*>(p.out + o + 8) = v1;
            } else {
                // one full 32-byte sector per thread and instruction (STG.256): the rows of a warp's lanes are Cout * 2 bytes
                // apart, so every 16-byte store was a partial-sector write request of its own
                asm volatile("st.global.v8.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};" ::"l"(p.out + o), "r"(v0.x), "r"(v0.y), "r"(v0.z), "r"(v0.w),
                             "r"(v1.x), "r"(v1.y), "r"(v1.z), "r"(v1.w) : "memory");
            }
        };

        pdl_wait();                          // residual reads and output writes follow the previous launch
        if (team < T) {
            locate();
        }
        int sr = team % p.a_stages, lap_r = team / p.a_stages;      // ring stage / lap of tile j
        const int groups = (MT * p.n_tile + 31) >> 5;               // 32-column groups of the accumulator buffer
        for (int j = team; j < T; j += teams) {
            const int abuf = j % p.acc_bufs, k = j / p.acc_bufs;
            const uint32_t t_lane = t_lane0 + abuf * acc_stride;
            const bool more = j + teams < T;
            const uint32_t rst = r_base + sr * r_tile_bytes;
            if (dbg && grp == 2 && lane == 0 && j < 16) dbg[160 + 4 * j] = clock64();           // team starts waiting for tile j
            if (res_smem) mbar_wait(res_full + 8u * sr, (uint32_t)(lap_r & 1));
            mbar_wait(acc_full + 8u * abuf, (uint32_t)(k & 1));
            if (dbg && warp == 4 && lane == 0 && j == 0) dbg[5] = clock64();
            if (dbg && grp == 2 && lane == 0 && j < 16) dbg[96 + 2 * j] = clock64();             // accumulator ready
            tc_fence_after();
            if (p.dbg_flags & 8) {            // bring-up: barrier handshakes only (measures the TMA + MMA side alone)
                tc_fence_before();
                __syncwarp();
                if (lane == 0) asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(acc_empty + 8u * abuf) : "memory");
            }
            for (int g = 0; g < ((p.dbg_flags & 8) ? 0 : groups); ++g) {
                // 32 accumulator columns -> registers; after the last group the TMEM buffer goes back to the MMA warps
                const int col = g << 5;
                const int mt = MT == 1 ? 0 : (col >= p.n_tile ? 1 : 0);
                const int c0 = col - mt * p.n_tile;
                const bool two = c0 + 32 <= p.n_tile;
                uint32_t ra[16], rb[16];
                tmem_ld16(t_lane + (uint32_t)col, ra);
                if (two) tmem_ld16(t_lane + (uint32_t)(col + 16), rb);
                tmem_ld_wait();
                if (dbg && grp == 2 && lane == 0 && j < 16 && g == 0) dbg[161 + 4 * j] = clock64();   // first 32 columns in registers
                if (g == groups - 1) {
                    tc_fence_before();
                    __syncwarp();
                    if (lane == 0) asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(acc_empty + 8u * abuf) : "memory");      // one arrival per epilogue warp
                }
                uint4 q[4] = {};
                const bool ok = mt ? valid[1] : valid[0];
                const size_t ob = mt ? obase[1] : obase[0];
                if (res_smem) {
                    res_lds(rst, mt, c0, q[0], q[1]);
                    if (two) res_lds(rst, mt, c0 + 16, q[2], q[3]);
                } else if (has_res && ok) {                 // residual straight from global memory (shapes the ring does not cover)
                    const uint4* rp = reinterpret_cast<const uint4*>(p.res + ob + c0);
                    q[0] = rp[0]; q[1] = rp[1];
                    if (two) { q[2] = rp[2]; q[3] = rp[3]; }
                }
                if (ok) {
                    finish16(ra, ob + c0, c0, q[0], q[1]);
                    if (two) finish16(rb, ob + c0 + 16, c0 + 16, q[2], q[3]);
                }
            }
            if (dbg && grp == 2 && lane == 0 && j < 16) dbg[162 + 4 * j] = clock64();           // all stores issued
            if (res_smem) {
                // residual stage back to the producer
                __syncwarp();
                if (lane == 0) asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(res_empty + 8u * sr) : "memory");
            }
            if (dbg && warp == 4 && lane == 0 && j == 0) dbg[6] = clock64();
            if (dbg && grp == 2 && lane == 0 && j < 16) dbg[97 + 2 * j] = clock64();             // epilogue done
            // residual of this team's next tile: in flight while the team waits for that accumulator
            if (more) { advance(); locate(); }
            sr += teams; while (sr >= p.a_stages) { sr -= p.a_stages; ++lap_r; }
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 1) {
        tc_fence_after();
        tmem_dealloc(tmem_base, p.tmem_cols);
        if (dbg && lane == 0) dbg[7] = clock64();
    }
    if (p.tl && threadIdx.x == 0) atomicMax(p.tl + 1, globaltimer_ns());
}


// ---------------------------------------------------------------------------
// Chain kernel: the 3x3 stride-1 convolutions of one resolution branch of a stage module (4 BasicBlocks = 8
// convolutions, same tensor shape throughout) as ONE persistent launch.  A kernel boundary between two dependent
// convolutions costs ~6.7 us at batch 64 (launch gap, prologue, first halo tile / weights, drain of the last tile:
// profiles/r02_branch_conv_slope.md) around 6 us of steady-state work -- but images are independent, so layer l+1 of
// a tile only needs layer l of the SAME image's neighbouring tiles.  Here every CTA owns a contiguous range of tiles
// (the same range in every layer) and walks layer 0, layer 1, ... over it; a tile's halo load waits for per-tile
// completion flags of the (up to nine) tiles of the previous layer it reads, written by whichever CTA owns them:
//   epilogue warp after its stores:  fence.proxy.async (generic writes -> TMA reads) ; __threadfence ; red.release.gpu +1
//   producer warp before the load :  ld.acquire.gpu until the flag shows 4 arrivals of this launch ; fence.proxy.async
// All CTAs of the launch are co-resident (grid <= SMs of the branch's share, one CTA per SM), dependencies only point
// to earlier layers, every CTA walks the layers in order: no deadlock.  Inside a CTA the walk starts at the first
// image boundary of its range, so that the first tiles of a layer read tiles the neighbour CTAs finished a whole
// layer earlier; the TMA ring, the TMEM double buffering and the epilogue teams run on across layer boundaries.
// Flags count arrivals over all launches (4 per tile and launch); the launch number comes from a device counter that
// the last CTA to leave increments, so graph replays need no memset node.
// Weights: resident set reloaded slot by slot at a layer switch (each issuer releases a slot after its last tile of
// the layer), or the streaming ring, which simply runs on into the next layer's weights.  Activation buffers may be
// reused two layers later (the program's own liveness pool does that): tile (l+2, t) starts only after every tile
// that read region t of the old contents has finished.
constexpr int kMaxChain = 8;
struct alignas(64) ChainEntry {
    CUtensorMap tmA, tmB, tmR;
    const float* bias;
    __half* out;
    int relu, has_res;
    int res_layer;               // layer of this chain whose output is the residual (-1: a tensor written before the launch)
    unsigned long long* tl;      // HBP_TIMELINE stamps of this member op
};
struct ChainHeader {
    int n_layers;
    unsigned* flags;             // [n_layers][n_tiles] arrival counters
    unsigned* state;             // [0] launches completed, [1] CTAs that have left the current launch
    unsigned long long* trace;   // bring-up (HBP_CHAIN_TRACE=<op name substring>): [cta][layer][first tile start, last tile end] globaltimer ns
};

__device__ __forceinline__ unsigned ld_acquire_gpu(const unsigned* p) {
    unsigned v;
    asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}

constexpr int kMaxChainTiles = 512;          // tiles per CTA and layer the chain kernel's shared tables hold
struct ChainTile {                           // one tile of this CTA's walk, by position
    uint32_t tile;                           // global tile index
    uint32_t origin;                         // element offset of the tile's first output pixel, channel 0
    uint16_t n0, tg;                         // first image, image-group index
    uint8_t tw, th, outer, pad;              // tile column / row; outer: some other CTA's halo reads this tile
};
struct ChainLayer {
    __half* out;
    unsigned long long* tl;
    int relu, has_res, res_layer;
};

template <int KSTEPS, int MT>
__global__ void __launch_bounds__(kHaloThreads, 1)
conv_umma_chain_kernel(const ChainEntry* __restrict__ table, const ConvParams p, const ChainHeader hdr) {
    constexpr int KS = 3, HW = kHaloW;
    extern __shared__ __align__(1024) uint8_t smem_raw[];
    __shared__ ChainTile s_tile[kMaxChainTiles];           // by walk position q
    __shared__ unsigned s_done[kMaxChainTiles];            // by local tile index: arrivals (4 per finished layer) of this launch
    __shared__ unsigned short s_pos[kMaxChainTiles];       // scratch of the order computation
    __shared__ ChainLayer s_layer[kMaxChain];
    const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
    const uint32_t a_tile_bytes = p.n_chunks * p.a_chunk_bytes;
    const uint32_t a_base = smem_base;
    const uint32_t b_base = a_base + p.a_stages * a_tile_bytes;
    const uint32_t r_base = b_base + p.b_slots * (uint32_t)KS * p.b_stage_bytes;
    const uint32_t r_tile_bytes = p.res_smem ? p.r_chunks * p.r_chunk_bytes : 0u;
    const uint32_t bar_base = r_base + p.a_stages * r_tile_bytes;
    const uint32_t a_full = bar_base;                                   // a_stages x n_chunks
    const uint32_t a_empty = a_full + 8u * (kMaxAStages * kMaxChunks);  // a_stages
    const uint32_t b_full = a_empty + 8u * kMaxAStages;                 // b_slots
    const uint32_t b_empty = b_full + 8u * kMaxBSlots;                  // b_slots
    const uint32_t acc_full = b_empty + 8u * kMaxBSlots;                // acc_bufs
    const uint32_t acc_empty = acc_full + 8u * kMaxAccBufs;             // acc_bufs
    const uint32_t res_full = acc_empty + 8u * kMaxAccBufs;             // a_stages
    const uint32_t res_empty = res_full + 8u * kMaxAStages;             // a_stages
    const uint32_t tmem_slot = res_empty + 8u * kMaxAStages;
    float* s_bias = reinterpret_cast<float*>(smem_raw + (bar_base + kHaloBarBytes - smem_u32(smem_raw)));   // [layer][n_tile]

    const int warp = uniform_warp_idx(), lane = threadIdx.x & 31;
    const int L = hdr.n_layers;
    const int tiles_per_img = p.tiles_w * p.tiles_h;
    // this CTA's tiles: [t0, t0 + T) in every layer
    const int t0 = (int)(((long long)blockIdx.x * p.n_tiles) / (long long)gridDim.x);
    const int T = (int)(((long long)(blockIdx.x + 1) * p.n_tiles) / (long long)gridDim.x) - t0;
    const int total = L * T;                                             // tile sequence numbers s = l * T + q
    const int n_slots = KS * p.n_chunks;                                 // weight slots per layer: (chunk, dx), three taps each

    pdl_launch_dependents();
    // Walk order inside a layer: first the tiles some OTHER CTA's halo reads (a neighbour tile outside [t0, t0+T)), then the
    // rest in index order.  What a neighbour CTA needs of layer l is then published early in its pass over layer l, and what
    // this CTA needs of its own interior was finished most of a layer ago: no CTA waits at a layer boundary.
    for (int i = threadIdx.x; i < T; i += kHaloThreads) {
        const int tile = t0 + i;
        const int tw = tile % p.tiles_w, th_ = (tile / p.tiles_w) % p.tiles_h, tg = tile / tiles_per_img;
        bool outer = false;
        for (int dh = -1; dh <= 1; ++dh)
            for (int dw = -1; dw <= 1; ++dw) {
                const int nh = th_ + dh, nw = tw + dw;
                if (nh < 0 || nh >= p.tiles_h || nw < 0 || nw >= p.tiles_w) continue;
                const int nt = (tg * p.tiles_h + nh) * p.tiles_w + nw;
                outer = outer || nt < t0 || nt >= t0 + T;
            }
        if (p.dbg_flags & 32) outer = false;
        s_pos[i] = outer ? 1 : 0;
        s_done[i] = 0u;
    }
    if ((int)threadIdx.x < L) {
        const ChainEntry& e = table[threadIdx.x];
        ChainLayer li;
        li.out = e.out; li.tl = e.tl; li.relu = e.relu; li.has_res = e.has_res; li.res_layer = e.res_layer;
        s_layer[threadIdx.x] = li;
    }
    __syncthreads();
    if (threadIdx.x == 0) {                  // stable partition (T is a few dozen): s_pos[i] <- walk position of local tile i
        int n_outer = 0;
        for (int i = 0; i < T; ++i) n_outer += s_pos[i];
        int a = 0, b = n_outer;
        for (int i = 0; i < T; ++i) { const bool o = s_pos[i] != 0; s_pos[i] = (unsigned short)((o ? a++ : b++) | (o ? 0x8000 : 0)); }
    }
    __syncthreads();
    for (int i = threadIdx.x; i < T; i += kHaloThreads) {
        const int tile = t0 + i;
        const int tw = tile % p.tiles_w, th_ = (tile / p.tiles_w) % p.tiles_h, tg = tile / tiles_per_img;
        ChainTile ct;
        ct.tile = (uint32_t)tile;
        ct.n0 = (uint16_t)(tg * p.tn); ct.tg = (uint16_t)tg;
        ct.tw = (uint8_t)tw; ct.th = (uint8_t)th_; ct.outer = (s_pos[i] & 0x8000) ? 1 : 0; ct.pad = 0;
        ct.origin = (uint32_t)((((size_t)tg * p.tn * p.Ho + (size_t)th_ * p.th) * p.Wo + (size_t)tw * 8) * p.Cout);
        s_tile[s_pos[i] & 0x7fff] = ct;
    }
    if (warp == 0) {
        for (int i = lane; i < L; i += 32) {
            asm volatile("prefetch.tensormap [%0];" ::"l"(&table[i].tmA) : "memory");
            asm volatile("prefetch.tensormap [%0];" ::"l"(&table[i].tmB) : "memory");
            if (table[i].has_res) asm volatile("prefetch.tensormap [%0];" ::"l"(&table[i].tmR) : "memory");
        }
        for (int i = lane; i < p.a_stages * p.n_chunks; i += 32) mbar_init(a_full + 8u * i, 1);
        for (int i = lane; i < p.a_stages; i += 32) { mbar_init(a_empty + 8u * i, 1); mbar_init(res_full + 8u * i, 1); mbar_init(res_empty + 8u * i, 4); }
        for (int i = lane; i < p.b_slots; i += 32) { mbar_init(b_full + 8u * i, 1); mbar_init(b_empty + 8u * i, p.b_resident ? (uint32_t)p.issuers : 1u); }
        for (int i = lane; i < kMaxAccBufs; i += 32) { mbar_init(acc_full + 8u * i, 1); mbar_init(acc_empty + 8u * i, 4); }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 1) tmem_alloc(tmem_slot, p.tmem_cols);
    for (int i = threadIdx.x; i < L * p.n_tile; i += kHaloThreads) s_bias[i] = table[i / p.n_tile].bias[i % p.n_tile];
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    uint32_t tmem_base;
    asm volatile("ld.shared.b32 %0, [%1];" : "=r"(tmem_base) : "r"(tmem_slot));
    tmem_base = __shfl_sync(0xffffffffu, tmem_base, 0);
    const uint32_t acc_stride = (uint32_t)(MT * p.n_tile);
    const uint32_t b_bytes = (uint32_t)p.n_tile * p.row_bytes;
    const uint32_t b_slot_bytes = (uint32_t)KS * p.b_stage_bytes;
    // spin with a watchdog: a dependency that never arrives is a bug -- trap instead of hanging the GPU
    auto spin_guard = [&](unsigned& spins, long long& t_start) {
        __nanosleep(32);
        if ((++spins & 4095u) == 0u) {
            const long long now = clock64();
            if (t_start == 0) t_start = now;
            else if (g_watchdog_cycles > 0 && now - t_start > g_watchdog_cycles) __trap();
        }
    };

    if (warp == 0) {
        // ===== producer: halo tiles (after their dependencies) and weights =====
        auto load_slot = [&](int l, int i, int slot) {      // weight slot i = (chunk, dx) of layer l -> ring / resident slot `slot`
            const int cc = i / KS, dx = i % KS;
            if (elect_one()) {
                mbar_expect_tx(b_full + 8u * slot, (uint32_t)KS * b_bytes);
                for (int dy = 0; dy < KS; ++dy)
                    tma_load_2d(b_base + slot * b_slot_bytes + dy * p.b_stage_bytes, &table[l].tmB, b_full + 8u * slot, cc * p.chunk,
                                (dy * KS + dx) * p.Cout);
            }
            __syncwarp();
        };
        pdl_wait();
        const unsigned want = 4u * (*reinterpret_cast<volatile const unsigned*>(hdr.state) + 1u);   // arrivals a finished tile shows in THIS launch
        int bs = 0;                         // streaming ring position
        uint32_t bph = 0;
        for (int s = 0; s < total; ++s) {
            const int l = s / T, q = s - l * T;
            const ChainTile ct = s_tile[q];
            if (l > 0 && !(p.dbg_flags & 4)) {
                // the (up to nine) tiles of layer l-1 this halo reads: lanes 0..8 check one each -- this CTA's own tiles in
                // shared memory, the neighbours' through their global flags
                const int dh = lane / 3 - 1, dw = lane % 3 - 1;
                const int nh = (int)ct.th + dh, nw = (int)ct.tw + dw;
                const bool mine = lane < 9 && nh >= 0 && nh < p.tiles_h && nw >= 0 && nw < p.tiles_w;
                const int nt = ((int)ct.tg * p.tiles_h + (mine ? nh : 0)) * p.tiles_w + (mine ? nw : 0);
                const int ni = nt - t0;
                const bool local = ni >= 0 && ni < T;
                const unsigned* f = hdr.flags + (size_t)(l - 1) * p.n_tiles + nt;
                const unsigned need_local = 4u * (unsigned)l;
                unsigned spins = 0;
                long long t_start = 0;
                while (true) {
                    bool ok = true;
                    if (mine) ok = local ? (*reinterpret_cast<volatile unsigned*>(&s_done[ni]) >= need_local) : (ld_acquire_gpu(f) >= want);
                    if (__all_sync(0xffffffffu, ok)) break;
                    spin_guard(spins, t_start);
                }
                __threadfence_block();
                asm volatile("fence.proxy.async;" ::: "memory");   // observed generic-proxy writes -> the TMA reads below
            }
            const int sa = s % p.a_stages, lap = s / p.a_stages;
            if (lap > 0) mbar_wait(a_empty + 8u * sa, (uint32_t)((lap - 1) & 1));
            if (elect_one()) {
                const int n0 = ct.n0, h0 = (int)ct.th * p.th, w0 = (int)ct.tw * 8;
                for (int cc = 0; cc < p.n_chunks; ++cc) {
                    const uint32_t bar = a_full + 8u * (sa * p.n_chunks + cc);
                    mbar_expect_tx(bar, p.a_box_bytes);
                    tma_load_4d(a_base + sa * a_tile_bytes + cc * p.a_chunk_bytes, &table[l].tmA, bar, cc * p.chunk, w0 - 1, h0 - 1, n0);
                }
            }
            __syncwarp();
            if (!p.b_resident) {
                // streaming ring: the weights of tile s, slot by slot (runs ahead of the MMA warp by b_slots slots)
                for (int i = 0; i < n_slots; ++i) {
                    mbar_wait(b_empty + 8u * bs, bph ^ 1u);
                    load_slot(l, i, bs);
                    if (++bs == p.b_slots) { bs = 0; bph ^= 1u; }
                }
            }
        }
    } else if (warp == 3) {
        // ===== residual producer: one ring stage per tile of EVERY layer (layers without a residual complete the stage empty),
        // so that the ring's phases stay aligned with the tile sequence =====
        // Resident weights are this warp's job too (the halo producer must never block on them: it runs ahead across the
        // layer boundary, so the halo ring is full when the new weights land).  Layer lt uses set lt % b_sets; its slots are
        // free once every issuer has passed its last tile of layer lt - b_sets (each commits b_empty there).
        auto load_wslot = [&](int l, int i) {
            const int cc = i / KS, dx = i % KS, slot = (l % p.b_sets) * n_slots + i;
            if (elect_one()) {
                mbar_expect_tx(b_full + 8u * slot, (uint32_t)KS * b_bytes);
                for (int dy = 0; dy < KS; ++dy)
                    tma_load_2d(b_base + slot * b_slot_bytes + dy * p.b_stage_bytes, &table[l].tmB, b_full + 8u * slot, cc * p.chunk,
                                (dy * KS + dx) * p.Cout);
            }
            __syncwarp();
        };
        if (p.b_resident)
            for (int l = 0; l < p.b_sets && l < L; ++l)
                for (int i = 0; i < n_slots; ++i) load_wslot(l, i);           // weights do not depend on the previous launch
        if (p.res_smem || p.b_resident) {
            pdl_wait();
            for (int s = 0; s < total; ++s) {
                const int l = s / T, q = s - l * T;
                if (p.b_resident && q == 0 && l > 0) {
                    const int lt = l + p.b_sets - 1;
                    if (lt < L)
                        for (int i = 0; i < n_slots; ++i) {
                            mbar_wait(b_empty + 8u * ((lt % p.b_sets) * n_slots + i), (uint32_t)((lt / p.b_sets - 1) & 1));
                            load_wslot(lt, i);
                        }
                }
                if (!p.res_smem) continue;
                const ChainTile ct = s_tile[q];
                const ChainLayer li = s_layer[l];
                const int sr = s % p.a_stages, lap = s / p.a_stages;
                if (lap > 0) mbar_wait(res_empty + 8u * sr, (uint32_t)((lap - 1) & 1));
                // the residual of layer l is the output of an earlier layer of this chain (same tile: written by this CTA's own
                // epilogue, normally a whole layer ago) or a tensor from before the launch
                if (li.has_res && li.res_layer >= 0 && !(p.dbg_flags & 4)) {
                    const unsigned need = 4u * (unsigned)(li.res_layer + 1);
                    const volatile unsigned* d = &s_done[(int)ct.tile - t0];
                    unsigned spins = 0;
                    long long t_start = 0;
                    while (*d < need) spin_guard(spins, t_start);
                    __threadfence_block();
                    asm volatile("fence.proxy.async;" ::: "memory");
                }
                if (elect_one()) {
                    if (li.has_res) {
                        mbar_expect_tx(res_full + 8u * sr, (uint32_t)p.r_chunks * p.r_box_bytes);
                        for (int rc = 0; rc < p.r_chunks; ++rc)
                            tma_load_4d(r_base + sr * r_tile_bytes + rc * p.r_chunk_bytes, &table[l].tmR, res_full + 8u * sr, rc * 64,
                                        (int)ct.tw * 8, (int)ct.th * p.th, ct.n0);
                    } else {
                        asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(res_full + 8u * sr) : "memory");
                    }
                }
                __syncwarp();
            }
        }
    } else if (warp <= 2) {
      if (warp - 1 < p.issuers) {
        // ===== MMA issuers: tile s by warp s % issuers into accumulator buffer s % acc_bufs =====
        const int iss = warp - 1;
        constexpr uint32_t kRow16 = KSTEPS * 2;
        constexpr uint32_t kMtStep = 16u * HW * kRow16;
        const uint64_t da0 = make_desc(0, p.row_bytes, HW * p.row_bytes);
        const uint64_t db0 = make_desc(0, p.row_bytes);
        const uint32_t bstep = p.b_stage_bytes >> 4;
        const uint32_t idesc = p.idesc, n_tile = (uint32_t)p.n_tile;
        int bs = 0;                         // streaming ring position (one issuer in that mode)
        uint32_t bph = 0;
        int seen_layer = -1;                // resident weights: layer whose slots this warp has already waited for
        for (int s = iss; s < total; s += p.issuers) {
            const int l = s / T;
            const int sa = s % p.a_stages, lap_a = s / p.a_stages;
            const int ab = s % p.acc_bufs, use = s / p.acc_bufs;
            const uint32_t pa = (uint32_t)(lap_a & 1);
            if (use > 0) {
                mbar_wait(acc_empty + 8u * ab, (uint32_t)((use - 1) & 1));
                tc_fence_after();
            }
            const uint32_t d_base = tmem_base + ab * acc_stride;
            const bool last_of_layer = p.b_resident && s + p.issuers >= (l + 1) * T && l + p.b_sets < L;   // my last tile of layer l: release its slots to layer l + b_sets
            const bool fast = p.b_resident && seen_layer == l && !last_of_layer;
            const int set0 = p.b_resident ? (l % p.b_sets) * n_slots : 0;
            for (int cc = 0; cc < p.n_chunks; ++cc) {
                mbar_wait(a_full + 8u * (sa * p.n_chunks + cc), pa);
                const uint64_t a_c = da0 + ((a_base + sa * a_tile_bytes + cc * p.a_chunk_bytes) >> 4);
                if (fast) {
                    tc_fence_after();
                    const uint64_t b_c = db0 + ((b_base + (uint32_t)(set0 + cc * KS) * b_slot_bytes) >> 4);
                    const uint32_t fresh = cc == 0 ? 0u : 1u;
                    if (elect_one()) {
#pragma unroll
                        for (int dx = 0; dx < KS; ++dx)
#pragma unroll
                            for (int dy = 0; dy < KS; ++dy)
#pragma unroll
                                for (int ks = 0; ks < KSTEPS; ++ks)
#pragma unroll
                                    for (int mt = 0; mt < MT; ++mt)
                                        umma_f16(d_base + mt * n_tile, a_c + (uint32_t)((dy * HW + dx) * kRow16 + mt * kMtStep + 2 * ks),
                                                 b_c + (uint32_t)(dx * KS + dy) * bstep + 2 * ks, idesc, (dx | dy | ks) ? 1u : fresh);
                    }
                    __syncwarp();
                    continue;
                }
#pragma unroll
                for (int dx = 0; dx < KS; ++dx) {
                    const int slot = p.b_resident ? set0 + cc * KS + dx : bs;
                    if (!p.b_resident) mbar_wait(b_full + 8u * bs, bph);
                    else if (seen_layer != l) mbar_wait(b_full + 8u * slot, (uint32_t)((l / p.b_sets) & 1));      // first tile of this layer for this warp
                    tc_fence_after();
                    const uint64_t b_s = db0 + ((b_base + slot * b_slot_bytes) >> 4);
                    const uint32_t fresh = (cc == 0 && dx == 0) ? 0u : 1u;
                    if (elect_one()) {
#pragma unroll
                        for (int dy = 0; dy < KS; ++dy) {
                            const uint64_t bd = b_s + (uint32_t)dy * bstep;
#pragma unroll
                            for (int ks = 0; ks < KSTEPS; ++ks)
#pragma unroll
                                for (int mt = 0; mt < MT; ++mt)
                                    umma_f16(d_base + mt * n_tile, a_c + (uint32_t)((dy * HW + dx) * kRow16 + mt * kMtStep + 2 * ks),
                                             bd + 2 * ks, idesc, (dy | ks) ? 1u : fresh);
                        }
                        if (!p.b_resident || last_of_layer) umma_commit(b_empty + 8u * slot);
                    }
                    __syncwarp();
                    if (!p.b_resident && ++bs == p.b_slots) { bs = 0; bph ^= 1u; }
                }
            }
            seen_layer = l;
            if (elect_one()) {
                umma_commit(a_empty + 8u * sa);
                umma_commit(acc_full + 8u * ab);
            }
            __syncwarp();
        }
      }
    } else {
        // ===== epilogue teams: tile s by team s % teams =====
        const int team = (warp - 4) >> 2;
        const int grp = warp & 3;
        const int r = grp * 32 + lane;
        const bool res_ring = p.res_smem != 0;
        const uint32_t t_lane0 = tmem_base + ((uint32_t)(grp * 32) << 16);
        const int teams = p.teams;
        bool row_ok[2];
        int row_n[2], row_w;
        uint32_t row_off[2];
        uint32_t rs_off[2], rs_xor[2];
        row_w = r & 7;
#pragma unroll
        for (int mt = 0; mt < 2; ++mt) {
            const int g = mt * 16 + (r >> 3);
            const int nn = g / p.rs, hh = g - nn * p.rs;
            row_ok[mt] = mt < MT && nn < p.tn && hh < p.th;
            row_n[mt] = nn;
            row_off[mt] = (uint32_t)((((size_t)nn * p.Ho + hh) * p.Wo + row_w) * p.Cout);
            const uint32_t rrow = row_ok[mt] ? (uint32_t)((nn * p.th + hh) * 8 + row_w) : 0u;
            rs_off[mt] = rrow * p.r_row_bytes;
            rs_xor[mt] = p.r_row_bytes == 128 ? (rrow & 7u) : ((rrow >> 1) & 3u);
        }
        auto res_lds = [&](uint32_t rst, int mt, int c0, uint4& q0, uint4& q1) {
            const uint32_t rso = mt ? rs_off[1] : rs_off[0], rsx = mt ? rs_xor[1] : rs_xor[0];
            const uint32_t box = rst + (uint32_t)(c0 >> 6) * p.r_chunk_bytes + rso;
            const uint32_t ci = (uint32_t)(c0 & 63) >> 3;
            const uint32_t a0 = box + (((ci) ^ rsx) << 4), a1 = box + (((ci + 1) ^ rsx) << 4);
            asm volatile("ld.shared.v4.b32 {%0, %1, %2, %3}, [%4];" : "=r"(q0.x), "=r"(q0.y), "=r"(q0.z), "=r"(q0.w) : "r"(a0));
            asm volatile("ld.shared.v4.b32 {%0, %1, %2, %3}, [%4];" : "=r"(q1.x), "=r"(q1.y), "=r"(q1.z), "=r"(q1.w) : "r"(a1));
        };
        pdl_wait();
        const int groups = (MT * p.n_tile + 31) >> 5;
        // A finished tile is published -- fence for the async proxy, +1 on its shared-memory counter, and for tiles another
        // CTA reads a device-scope fence and +1 on the global flag -- just BEFORE the stores of this team's next tile (its own
        // stores were issued a tile period earlier and have been acknowledged by then, so the fences do not wait for them), or
        // at once when the team would otherwise sit waiting for its next accumulator.
        unsigned* pend_g = nullptr;
        int pend_i = -1;
        auto publish = [&]() {
            if (!(p.dbg_flags & 2)) asm volatile("fence.proxy.async;" ::: "memory");
            if (pend_g && !(p.dbg_flags & 1)) __threadfence(); else __threadfence_block();
            __syncwarp();
            if (lane == 0) {
                if (pend_g) asm volatile("red.release.gpu.global.add.u32 [%0], 1;" ::"l"(pend_g) : "memory");
                atomicAdd(&s_done[pend_i], 1u);
            }
            pend_g = nullptr; pend_i = -1;
        };
        for (int s = team; team < teams && s < total; s += teams) {
            const int l = s / T, q = s - l * T;
            const ChainTile ct = s_tile[q];
            const ChainLayer li = s_layer[l];
            const bool valid0 = row_ok[0] && (int)ct.n0 + row_n[0] < p.P && (int)ct.tw * 8 + row_w < p.Wo;
            const bool valid1 = row_ok[1] && (int)ct.n0 + row_n[1] < p.P && (int)ct.tw * 8 + row_w < p.Wo;
            __half* const out0 = li.out + (size_t)(ct.origin + row_off[0]);
            __half* const out1 = li.out + (size_t)(ct.origin + row_off[1]);
            const bool has_res = li.has_res != 0, relu = li.relu != 0;
            const float* bias_l = s_bias + l * p.n_tile;
            const int abuf = s % p.acc_bufs, k = s / p.acc_bufs;
            const int sr = s % p.a_stages, lap_r = s / p.a_stages;
            const uint32_t t_lane = t_lane0 + abuf * acc_stride;
            const uint32_t rst = r_base + sr * r_tile_bytes;
            if (li.tl && q == 0 && grp == 0 && lane == 0) atomicMin(li.tl, globaltimer_ns());
            if (hdr.trace && q == 0 && grp == 0 && lane == 0) hdr.trace[((size_t)blockIdx.x * L + l) * 2] = globaltimer_ns();
            if (pend_i >= 0) {
                // next accumulator not there yet: nothing better to do than to publish now
                uint32_t ready;
                asm volatile(
                    "{\n"
                    ".reg .pred p;\n"
                    "mbarrier.test_wait.parity.shared::cta.b64 p, [%1], %2;\n"
                    "selp.u32 %0, 1, 0, p;\n"
                    "}\n" : "=r"(ready) : "r"(acc_full + 8u * abuf), "r"((uint32_t)(k & 1)) : "memory");
                if (!__all_sync(0xffffffffu, ready != 0)) publish();
            }
            if (res_ring) mbar_wait(res_full + 8u * sr, (uint32_t)(lap_r & 1));
            mbar_wait(acc_full + 8u * abuf, (uint32_t)(k & 1));
            tc_fence_after();
            for (int g = 0; g < groups; ++g) {
                const int col = g << 5;
                const int mt = MT == 1 ? 0 : (col >= p.n_tile ? 1 : 0);
                const int c0 = col - mt * p.n_tile;
                const bool two = c0 + 32 <= p.n_tile;
                uint32_t ra[16], rb[16];
                tmem_ld16(t_lane + (uint32_t)col, ra);
                if (two) tmem_ld16(t_lane + (uint32_t)(col + 16), rb);
                tmem_ld_wait();
                if (g == groups - 1) {
                    tc_fence_before();
                    __syncwarp();
                    if (lane == 0) asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(acc_empty + 8u * abuf) : "memory");
                }
                const bool ok = mt ? valid1 : valid0;
                __half* const o = mt ? out1 : out0;
                uint4 qv[4] = {};
                if (has_res) {
                    res_lds(rst, mt, c0, qv[0], qv[1]);
                    if (two) res_lds(rst, mt, c0 + 16, qv[2], qv[3]);
                }
                if (pend_i >= 0) publish();          // (before this tile's first store)
                if (ok) {
#pragma unroll
                    for (int hf = 0; hf < 2; ++hf) {
                        if (hf == 1 && !two) break;
                        const uint32_t* rr = hf ? rb : ra;
                        const int cc0 = c0 + 16 * hf;
                        float x[16];
#pragma unroll
                        for (int q4 = 0; q4 < 4; ++q4) {
                            const float4 b4 = *reinterpret_cast<const float4*>(bias_l + cc0 + 4 * q4);
                            x[4 * q4] = __uint_as_float(rr[4 * q4]) + b4.x; x[4 * q4 + 1] = __uint_as_float(rr[4 * q4 + 1]) + b4.y;
                            x[4 * q4 + 2] = __uint_as_float(rr[4 * q4 + 2]) + b4.z; x[4 * q4 + 3] = __uint_as_float(rr[4 * q4 + 3]) + b4.w;
                        }
                        if (has_res) {
                            const __half2* h0p = reinterpret_cast<const __half2*>(&qv[2 * hf]);
                            const __half2* h1p = reinterpret_cast<const __half2*>(&qv[2 * hf + 1]);
#pragma unroll
                            for (int u = 0; u < 4; ++u) {
                                const float2 f0 = __half22float2(h0p[u]), f1 = __half22float2(h1p[u]);
                                x[2 * u] += f0.x; x[2 * u + 1] += f0.y;
                                x[8 + 2 * u] += f1.x; x[8 + 2 * u + 1] += f1.y;
                            }
                        }
                        if (relu) {
#pragma unroll
                            for (int u = 0; u < 16; ++u) x[u] = fmaxf(x[u], 0.f);
                        }
                        __align__(16) __half2 pk[8];
#pragma unroll
                        for (int u = 0; u < 8; ++u) pk[u] = __floats2half2_rn(x[2 * u], x[2 * u + 1]);
                        store_half16(o + cc0, pk);
                    }
                }
            }
            if (res_ring) {
                __syncwarp();
                if (lane == 0) asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(res_empty + 8u * sr) : "memory");
            }
            // this warp's rows are stored; the fourth arrival completes the tile
            if (l + 1 < L) {
                pend_i = (int)ct.tile - t0;
                pend_g = ct.outer ? hdr.flags + (size_t)l * p.n_tiles + ct.tile : nullptr;
            }
            if (li.tl && q == T - 1 && grp == 0 && lane == 0) atomicMax(li.tl + 1, globaltimer_ns());
            if (hdr.trace && q == T - 1 && grp == 0 && lane == 0) hdr.trace[((size_t)blockIdx.x * L + l) * 2 + 1] = globaltimer_ns();
        }
        if (pend_i >= 0) publish();
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 1) {
        tc_fence_after();
        tmem_dealloc(tmem_base, p.tmem_cols);
    }
    if (threadIdx.x == 0) {
        // last CTA out closes the launch: the next launch's tiles then show 4 more arrivals
        __threadfence();
        const unsigned left = atomicAdd(hdr.state + 1, 1u);
        if (left == gridDim.x - 1) {
            hdr.state[1] = 0u;
            __threadfence();
            atomicAdd(hdr.state, 1u);
        }
    }
}

int halo_mode_enabled() {
    static int v = -1;
    if (v < 0) { const char* e = getenv("HBP_CONV_HALO"); v = e ? atoi(e) : 1; }
    return v;
}

bool pick_tile(int Ho, int Wo, int m_tiles, int* tn, int* th, int* tw) {
    const int rows = 128 * m_tiles;
    for (int n = 1; n <= 32; n *= 2) {
        if (rows % n) continue;
        const int per = rows / n;
        for (int w = Wo < per ? Wo : per; w >= 1; --w) {
            if (Wo % w || per % w) continue;
            const int h = per / w;
            if (h > Ho || Ho % h) continue;
            if (w > 256 || h > 256) continue;
            *tn = n; *th = h; *tw = w;
            return true;
        }
    }
    return false;
}

int stride_mode() {
    // how a tiled tensor map with elementStrides = s sizes its box:
    //   1 (default): boxDim counts SOURCE elements spanned, ceil(box/s) are written
    //   2          : boxDim counts elements written
    //   0          : stride-2 convolutions stay on the SIMT engine
    static int mode = -1;
    if (mode < 0) {
        const char* e = getenv("HBP_TMA_STRIDE_MODE");
        mode = e ? atoi(e) : 1;
    }
    return mode;
}

}  // namespace

bool umma_supported(const HrnetModel& m, const HOp& op) {
    if (op.kind != OP_CONV) return false;
    if (!(op.k == 1 || op.k == 3) || !(op.stride == 1 || op.stride == 2)) return false;
    if (op.stride == 2 && stride_mode() == 0) return false;
    if (!(op.cin == 32 || op.cin % 64 == 0)) return false;
    if (op.cout % 16 != 0) return false;
    const HTensor& ti = m.tensors[op.in];
    int tn, th, tw;
    if (!pick_tile(ti.h / op.stride, ti.w / op.stride, 1, &tn, &th, &tw)) return false;
    return get_encode() != nullptr;
}

void umma_plan_destroy(UmmaPlan* p) { delete p; }

static int encode_weights_map(EncodeTiledFn enc, UmmaPlan* pl, const HrnetModel& m, const HOp& op, int chunk,
                              int n_tile, CUtensorMapSwizzle sw) {
    cuuint64_t gdim[2] = {(cuuint64_t)op.cin, (cuuint64_t)op.k * op.k * op.cout};
    cuuint64_t gstr[1] = {(cuuint64_t)op.cin * 2};
    cuuint32_t box[2] = {(cuuint32_t)chunk, (cuuint32_t)n_tile};
    cuuint32_t est[2] = {1, 1};
    CUresult r = enc(&pl->tmB, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 2, m.d_weights + op.w_off, gdim, gstr, box, est,
                     CU_TENSOR_MAP_INTERLEAVE_NONE, sw, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                     CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) {
        hbp_set_error("cuTensorMapEncodeTiled(B) failed (%d) for %s", (int)r, op.name.c_str());
        return HBP_ERR_CUDA;
    }
    return HBP_OK;
}

static int env_int(const char* name, int dflt) {
    const char* e = getenv(name);
    return e ? atoi(e) : dflt;
}

// halo-mode plan (3x3, stride 1).  Returns HBP_OK with *ok = false when the shape does not fit.
static int plan_halo(hbp_ctx* ctx, HrnetModel& m, const HOp& op, int capP, UmmaPlan* pl, bool* ok) {
    // SMs this launch may occupy (one persistent CTA each): the whole GPU, or the share of its
    // resolution branch when the branches of a stage run side by side on their own streams
    int sm_budget = op.sm_share > 0.f ? (int)(op.sm_share * ctx->sm_count + 0.5f) : ctx->sm_count;
    if (env_int("HBP_HALO_SMS", 0)) sm_budget = env_int("HBP_HALO_SMS", 0);
    if (sm_budget < 1) sm_budget = 1;
    if (sm_budget > ctx->sm_count) sm_budget = ctx->sm_count;
    *ok = false;
    const HTensor& ti = m.tensors[op.in];
    const int Ho = ti.h, Wo = ti.w;
    ConvParams& p = pl->prm;
    const int ksz = op.k, hw = ksz == 3 ? kHaloW : 8;
    const int chunk = op.cin == 32 ? 32 : 64, n_chunks = op.cin / chunk;
    if (n_chunks > kMaxChunks) return HBP_OK;
    const uint32_t row_bytes = chunk * 2;
    const int tiles_w = (Wo + 7) / 8;
    const int sms = sm_budget;
    const uint32_t budget = (uint32_t)env_int("HBP_HALO_SMEM_KB", 216) * 1024u;   // (216: layer1's conv3 gets a four-stage halo + residual ring, 52 -> 39 us)

    // Candidate tiles: (tn images x th rows, m M-tiles).  One persistent CTA per SM walks
    // ceil(items / SMs) work items of 16*m stacked 8-pixel rows each, so the time of the launch
    // is ~ rounds * m * (MMA time of one M-tile) / (share of useful MMA rows is already inside
    // `items`): pick the (tile, N split) with the fewest M-tile rounds; ties go to the larger N
    // tile (fewer shared-memory bytes per MAC), then to the larger M (weights reused).
    const double l2_bpc = (double)env_int("HBP_HALO_L2BPC", 64);      // L2 -> shared memory bytes per clock and SM the cost model assumes
    struct Cand { int tn, th, m, n_tile; long items; double cost; };
    Cand best{0, 0, 0, 0, 0, 1e30};
    const int mt_max = (env_int("HBP_HALO_C64_M1", 0) && op.cin == 64 && op.cout == 64 && ksz == 3 && op.sm_share > 0.f && (env_int("HBP_HALO_C64_M1", 0) > 1 || op.res >= 0)) ? 1 : 2;
    for (int mt = 1; mt <= mt_max; ++mt)
        for (int n = 1; n <= 6; ++n)
            for (int h = Ho < 16 * mt ? Ho : 16 * mt; h >= 1; --h) {
                if (Ho % h) continue;
                if (n > 1 && h != Ho) continue;                 // stacked images need whole images
                const int rs = h + ksz - 1;
                if (16 * mt < n * rs - (ksz - 1)) continue;             // every group of the tile inside the M-tiles
                if (n * h * 2 < 16 * mt && !(n == 1 && h == Ho)) { break; }   // < 50 % useful rows: only if nothing else
                const long tiles = (long)((capP + n - 1) / n) * (Ho / h) * tiles_w;
                // (N = 64 tiles for layer1's wide 1x1 convs win 10 us per conv in isolation and lose 80 us per forward
                // in the network, where their tails overlap the next conv's prologue: HBP_HALO_1X1_NMAX, default off)
                const int c_top = (ksz == 1 && op.cout >= 256) ? env_int("HBP_HALO_1X1_NMAX", 256) : 256;
                for (int c = op.cout < c_top ? op.cout : c_top; c >= 32; c -= 16) {
                    if (op.cout % c || mt * c > 512) continue;
                    const long items = tiles * (op.cout / c);
                    const long rounds = (items + sms - 1) / sms;
                    // cycles of one M-tile pass over all taps: max(tensor floor, shared-memory feed)
                    const double per_mma = std::max(c / 2.0, (128.0 * 32 + c * 32.0) / 128.0);
                    const double mma_cyc = (double)rounds * mt * per_mma * ksz * ksz * (op.cin / 16);
                    // weight bytes one CTA pulls from L2 (~24 B/clk/SM when every SM streams): once when
                    // they can stay resident, once per tile otherwise
                    const double w_bytes = (double)ksz * ksz * op.cin * c * 2;
                    const double a_bytes = (double)(16 * mt + ksz - 1) * hw * op.cin * 2;      // one tile's input box
                    const bool fits = w_bytes + 2.0 * a_bytes < 190e3;
                    const double w_cyc = (fits ? 1.0 : (double)rounds) * w_bytes / l2_bpc;
                    // the tile's input box is fetched once per output-channel split: it bounds the 1x1 layers
                    const double a_cyc = (double)rounds * a_bytes / l2_bpc;
                    const double cost = std::max(std::max(mma_cyc, w_cyc), a_cyc);
                    const bool better = cost < best.cost * 0.97 ||
                                        (cost < best.cost * 1.03 && (c > best.n_tile || (c == best.n_tile && mt > best.m)));
                    if (better) best = Cand{n, h, mt, c, items, cost};
                }
                break;                                          // largest h for this (m, n)
            }
    if (!best.m) return HBP_OK;
    int tn = best.tn, th = best.th, m_tiles = best.m, n_tile = best.n_tile;
    if (env_int("HBP_HALO_M", 0)) {                             // bring-up override: force the M-tile count
        const int fm = env_int("HBP_HALO_M", 0);
        int ftn = 1, fth = 0, fmm = fm;
        (void)fmm;
        for (int h = Ho < 16 * fm ? Ho : 16 * fm; h >= 1; --h) if (Ho % h == 0) { fth = h; break; }
        if (fth) { tn = ftn; th = fth; m_tiles = fm; }
    }
    if (env_int("HBP_HALO_N", 0) && op.cout % env_int("HBP_HALO_N", 0) == 0) n_tile = env_int("HBP_HALO_N", 0);
    const int tiles_h = Ho / th;
    const int rs = th + ksz - 1;
    const long tiles = (long)((capP + tn - 1) / tn) * tiles_h * tiles_w;
    const int n_splits = op.cout / n_tile;
    const long per_cta = (tiles * n_splits + sms - 1) / sms;            // tiles one CTA walks

    const uint32_t a_box_bytes = (uint32_t)(hw * rs * tn) * row_bytes;
    uint32_t a_chunk_bytes = (uint32_t)((16 * m_tiles + ksz - 1) * hw) * row_bytes;
    if (a_chunk_bytes < a_box_bytes) a_chunk_bytes = a_box_bytes;
    a_chunk_bytes = (a_chunk_bytes + 1023u) & ~1023u;
    const uint32_t b_stage = ((uint32_t)n_tile * row_bytes + 1023u) & ~1023u;
    // residual through shared memory: one TMA box of 64 (or 32) channels x the tile's pixels per stage,
    // in the swizzle that makes the epilogue's 16-byte row reads conflict-free
    const int r_ch = n_tile < 64 ? n_tile : 64;
    const bool res_smem = op.res >= 0 && env_int("HBP_HALO_RES_SMEM", 1) && (r_ch == 32 || r_ch == 64) && n_tile % r_ch == 0;
    const int r_chunks = res_smem ? n_tile / r_ch : 0;
    const uint32_t r_row_bytes = (uint32_t)r_ch * 2;
    const uint32_t r_box_bytes = (uint32_t)(8 * th * tn) * r_row_bytes;
    const uint32_t r_chunk_bytes = (r_box_bytes + 1023u) & ~1023u;
    const uint32_t a_tile = a_chunk_bytes * n_chunks + (res_smem ? r_chunks * r_chunk_bytes : 0u);     // one ring stage: halo tile + residual tile
    // bias and residual through the tensor pipe (one extra MMA per M-tile / two per 32 output channels), so that the
    // epilogue issues no shared-memory loads.  Measured (profiles/r02_epilogue_ablation.md): 4.5 % on the 32-channel
    // branch conv, against 19 % from keeping the per-M-tile row geometry out of local memory; the tensor pipe does not
    // round like the epilogue's fp32 adds, so a conv would no longer give bit-identical results in the halo and the
    // per-tap kernels (plans change with the buffer capacity).  Off by default: HBP_HALO_BIAS_MMA=1 / HBP_HALO_RES_MMA=1.
    const bool bias_mma = env_int("HBP_HALO_BIAS_MMA", 0) != 0;
    const bool res_mma = res_smem && tn == 1 && n_tile % 32 == 0 && bias_mma && env_int("HBP_HALO_RES_MMA", 1) != 0;
    const uint32_t c_bytes = bias_mma ? 8192u + 2048u + (((uint32_t)n_tile * 64u + 1023u) & ~1023u) : 0u;
    const uint32_t fixed = kHaloBarBytes + (uint32_t)n_tile * 4 + 1024 + 64 + c_bytes;
    const int k_slots = ksz * n_chunks;                                 // weight slots per tile: (chunk, dx), ksz taps each
    const uint32_t b_slot = (uint32_t)ksz * b_stage;
    const uint32_t b_all = (uint32_t)k_slots * b_slot;
    int want_a = (int)std::min<long>(per_cta, env_int("HBP_HALO_ASTAGES", 4));
    if (want_a < 1) want_a = 1;
    if (want_a > kMaxAStages) want_a = kMaxAStages;
    // resident weights when they fit beside at least min(2, want_a) halo stages
    int a_stages = 0, b_slots = 0, resident = 0;
    if (env_int("HBP_HALO_RESIDENT", 1) && fixed + b_all + a_tile <= budget) {
        int fit = (int)((budget - fixed - b_all) / a_tile);
        if (fit >= std::min(env_int("HBP_HALO_RES_MINA", 2), want_a)) { resident = 1; a_stages = std::min(fit, want_a); b_slots = k_slots; }
    }
    if (!resident) {
        a_stages = std::min(want_a, 2);
        while (a_stages > 1 && fixed + a_stages * a_tile + 3 * b_slot > budget) --a_stages;
        if (fixed + a_stages * a_tile + 2 * b_slot > budget) return HBP_OK;
        b_slots = (int)((budget - fixed - a_stages * a_tile) / b_slot);
        if (b_slots > k_slots * (int)std::min<long>(per_cta, 2)) b_slots = k_slots * (int)std::min<long>(per_cta, 2);
        if (b_slots > kMaxBSlots) b_slots = kMaxBSlots;
        if (b_slots < 2) return HBP_OK;
    }
    if (a_stages == 3) a_stages = 2;         // even ring depth (see the note on parity waits below)
    // accumulator double-buffered in TMEM whenever two buffers fit
    int acc_bufs = 4 * m_tiles * n_tile <= 512 ? 4 : (2 * m_tiles * n_tile <= 512 ? 2 : 1);
    if (env_int("HBP_HALO_BUFS", 0)) acc_bufs = env_int("HBP_HALO_BUFS", 0);
    if (acc_bufs * m_tiles * n_tile > 512) return HBP_OK;
    if (rs > 256 || tn > 256) return HBP_OK;

    p.Ho = Ho; p.Wo = Wo; p.Cout = op.cout; p.up = 1; p.relu = op.relu;
    p.tn = tn; p.th = th; p.tw = 8; p.m_tiles = m_tiles; p.n_tile = n_tile;
    p.chunk = chunk; p.n_chunks = n_chunks; p.ksz = ksz; p.stride = 1;
    p.tiles_w = tiles_w; p.tiles_h = tiles_h;
    p.row_bytes = row_bytes;
    p.a_stage_bytes = 0; p.b_stage_bytes = b_stage; p.tx_bytes = 0;
    p.stages = b_slots;
    p.res_smem = res_smem ? 1 : 0; p.r_chunks = r_chunks; p.r_chunk_bytes = r_chunk_bytes; p.r_box_bytes = r_box_bytes;
    p.r_row_bytes = r_row_bytes;
    p.dbg_flags = env_int("HBP_HALO_DBG", 0);
    // Every mbarrier is waited on by parity, so a waiter must never run two phases ahead of it: ring
    // stage s and accumulator buffer b are always served by the same issuer warp / epilogue team,
    // i.e. the issuer and team counts divide the stage and buffer counts.
    p.teams = (acc_bufs % 2 == 0 && a_stages % 2 == 0) ? 2 : 1;
    // (a third epilogue team -- 512 threads, tile j -> team j % 3 -- gave no gain on the epilogue-heavy 1x1 convs, which are L2-byte bound, and
    // faulted intermittently when two forwards shared the GPU: removed, profiles/r02_epilogue_ablation.md)
    p.issuers = (resident && p.teams >= 2 && per_cta > 1) ? env_int("HBP_HALO_ISSUERS", 2) : 1;
    if (a_stages % p.issuers || acc_bufs % p.issuers) p.issuers = 1;
    p.acc_bufs = acc_bufs; p.a_stages = a_stages; p.b_slots = b_slots; p.b_resident = resident;
    uint32_t cols = 32;
    while (cols < (uint32_t)(acc_bufs * m_tiles * n_tile)) cols *= 2;
    p.tmem_cols = cols;
    p.idesc = (1u << 4) | ((uint32_t)(n_tile >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
    p.idesc32 = (1u << 4) | ((uint32_t)(32 >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
    p.bias_mma = bias_mma ? 1 : 0; p.res_mma = res_mma ? 1 : 0; p.c_bytes = c_bytes;
    p.mode = 1; p.rs = rs; p.a_chunk_bytes = a_chunk_bytes; p.a_box_bytes = a_box_bytes;
    {
        static const int l2_hints = env_int("HBP_L2_HINTS", 1), rev_ok = env_int("HBP_REVERSE", 1);
        p.reverse = (op.reverse && rev_ok) ? 1 : 0;
        p.hint_a = (op.in_dead && (l2_hints & 1)) ? kEvictFirst : kEvictNormal;
        p.hint_r = (op.res_dead && (l2_hints & 1)) ? kEvictFirst : kEvictNormal;
        p.hint_o = (op.out_keep && (l2_hints & 2)) ? kEvictLast : 0ull;
        p.split_producer = env_int("HBP_HALO_SPLIT_PRODUCER", 0);   // (transition1 alone 72 -> 65 us, the 128-channel branch 9.75 -> 9.46 us on 148 SMs; nothing inside the network: off)
    }
    p.bias = m.d_bias + op.b_off;
    p.res = op.res >= 0 ? m.bufs[m.tensors[op.res].buf] : nullptr;
    p.out = m.bufs[m.tensors[op.out].buf];
    pl->smem_bytes = (size_t)a_stages * a_tile + (size_t)b_slots * b_slot + fixed;
    pl->n_splits = n_splits;
    pl->occ = 1;
    pl->sm_budget = sm_budget;
    if (getenv("HBP_CONV_TRACE"))
        fprintf(stderr, "[plan] %s halo tile tn=%d th=%d m=%d n_tile=%d acc_bufs=%d a_stages=%d b_slots=%d resident=%d tmem=%u smem=%zu tiles=%ld per_cta=%ld sms=%d bias_mma=%d res_mma=%d\n",
                op.name.c_str(), tn, th, m_tiles, n_tile, acc_bufs, a_stages, b_slots, resident, cols, pl->smem_bytes, tiles, per_cta, sm_budget, p.bias_mma, p.res_mma);

    EncodeTiledFn enc = get_encode();
    const CUtensorMapSwizzle sw = row_bytes == 128 ? CU_TENSOR_MAP_SWIZZLE_128B : CU_TENSOR_MAP_SWIZZLE_64B;
    cuuint64_t gdim[4] = {(cuuint64_t)ti.c, (cuuint64_t)ti.w, (cuuint64_t)ti.h, (cuuint64_t)capP};
    cuuint64_t gstr[3] = {(cuuint64_t)ti.c * 2, (cuuint64_t)ti.w * ti.c * 2, (cuuint64_t)ti.h * ti.w * ti.c * 2};
    cuuint32_t box[4] = {(cuuint32_t)chunk, (cuuint32_t)hw, (cuuint32_t)rs, (cuuint32_t)tn};
    cuuint32_t est[4] = {1, 1, 1, 1};
    CUresult r = enc(&pl->tmA, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 4, m.bufs[ti.buf], gdim, gstr, box, est,
                     CU_TENSOR_MAP_INTERLEAVE_NONE, sw, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                     CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) {
        hbp_set_error("cuTensorMapEncodeTiled(A halo) failed (%d) for %s box=(%d,%d,%d,%d)", (int)r, op.name.c_str(),
                      chunk, hw, rs, tn);
        return HBP_ERR_CUDA;
    }
    pl->tmR = pl->tmA;                       // placeholder when the residual does not go through shared memory
    if (res_smem) {
        const HTensor& tr = m.tensors[op.res];
        cuuint64_t rdim[4] = {(cuuint64_t)tr.c, (cuuint64_t)tr.w, (cuuint64_t)tr.h, (cuuint64_t)capP};
        cuuint64_t rstr[3] = {(cuuint64_t)tr.c * 2, (cuuint64_t)tr.w * tr.c * 2, (cuuint64_t)tr.h * tr.w * tr.c * 2};
        cuuint32_t rbox[4] = {(cuuint32_t)r_ch, 8u, (cuuint32_t)th, (cuuint32_t)tn};
        CUresult rr = enc(&pl->tmR, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 4, m.bufs[tr.buf], rdim, rstr, rbox, est,
                          CU_TENSOR_MAP_INTERLEAVE_NONE, r_ch == 64 ? CU_TENSOR_MAP_SWIZZLE_128B : CU_TENSOR_MAP_SWIZZLE_64B,
                          CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        if (rr != CUDA_SUCCESS) {
            hbp_set_error("cuTensorMapEncodeTiled(residual) failed (%d) for %s box=(%d,8,%d,%d)", (int)rr, op.name.c_str(), r_ch, th, tn);
            return HBP_ERR_CUDA;
        }
    }
    int st = encode_weights_map(enc, pl, m, op, chunk, n_tile, sw);
    if (st) return st;
    *ok = true;
    return HBP_OK;
}

int umma_plan_create(hbp_ctx* ctx, HrnetModel& m, int op_index, int capP, UmmaPlan** out, bool for_group) {
    const HOp& op = m.ops[op_index];
    const bool persistent_tiling = for_group || op.persist || (env_int("HBP_PERSIST0", 2) > 1 && op.sm_share > 0.f);
    const HTensor& ti = m.tensors[op.in];
    const int Ho = ti.h / op.stride, Wo = ti.w / op.stride;
    UmmaPlan* pl = new UmmaPlan();
    ConvParams& p = pl->prm;
    memset(&p, 0, sizeof(p));
    p.reverse = (op.reverse && env_int("HBP_REVERSE", 1)) ? 1 : 0;
    EncodeTiledFn enc = get_encode();
    if (!enc) { delete pl; hbp_set_error("cuTensorMapEncodeTiled unavailable"); return HBP_ERR_CUDA; }
    if (!(ctx->attr_flags & ATTR_UMMA)) {
        {
            const char* e = getenv("HBP_WATCHDOG_S");
            const long long cycles = (long long)((e ? atof(e) : 20.0) * 2.0e9);
            HBP_CUDA(cudaMemcpyToSymbol(g_watchdog_cycles, &cycles, sizeof(cycles)));
        }
        HBP_CUDA(cudaFuncSetAttribute(conv_umma_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 220 * 1024));
        HBP_CUDA(cudaFuncSetAttribute(conv_umma_pgroup_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
        HBP_CUDA(cudaFuncSetAttribute(conv_umma_halo_kernel<2, 1, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize, 220 * 1024));
        HBP_CUDA(cudaFuncSetAttribute(conv_umma_halo_kernel<2, 1, 3>, cudaFuncAttributeMaxDynamicSharedMemorySize, 220 * 1024));
        HBP_CUDA(cudaFuncSetAttribute(conv_umma_halo_kernel<2, 2, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize, 220 * 1024));
        HBP_CUDA(cudaFuncSetAttribute(conv_umma_halo_kernel<2, 2, 3>, cudaFuncAttributeMaxDynamicSharedMemorySize, 220 * 1024));
        HBP_CUDA(cudaFuncSetAttribute(conv_umma_halo_kernel<4, 1, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize, 220 * 1024));
        HBP_CUDA(cudaFuncSetAttribute(conv_umma_halo_kernel<4, 1, 3>, cudaFuncAttributeMaxDynamicSharedMemorySize, 220 * 1024));
        HBP_CUDA(cudaFuncSetAttribute(conv_umma_halo_kernel<4, 2, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize, 220 * 1024));
        HBP_CUDA(cudaFuncSetAttribute(conv_umma_halo_kernel<4, 2, 3>, cudaFuncAttributeMaxDynamicSharedMemorySize, 220 * 1024));
        HBP_CUDA(cudaFuncSetAttribute(conv_umma_chain_kernel<2, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize, 212 * 1024));
        HBP_CUDA(cudaFuncSetAttribute(conv_umma_chain_kernel<2, 2>, cudaFuncAttributeMaxDynamicSharedMemorySize, 212 * 1024));
        HBP_CUDA(cudaFuncSetAttribute(conv_umma_chain_kernel<4, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize, 212 * 1024));
        HBP_CUDA(cudaFuncSetAttribute(conv_umma_chain_kernel<4, 2>, cudaFuncAttributeMaxDynamicSharedMemorySize, 212 * 1024));
        ctx->attr_flags |= ATTR_UMMA;
    }
    if (!for_group && (op.k == 3 || (op.k == 1 && env_int("HBP_HALO_1X1", 1))) && op.stride == 1 && op.up == 1 && halo_mode_enabled()) {
        bool ok = false;
        int st = plan_halo(ctx, m, op, capP, pl, &ok);
        if (st) { delete pl; return st; }
        if (ok) { *out = pl; return HBP_OK; }
        memset(&p, 0, sizeof(p));
    }
    // N tile: largest divisor of Cout that is a multiple of 16 and <= 256
    int n_tile = 0;
    for (int c = op.cout < 256 ? op.cout : 256; c >= 16; c -= 16)
        if (op.cout % c == 0) { n_tile = c; break; }
    if (!n_tile) { delete pl; hbp_set_error("no N tile for Cout=%d", op.cout); return HBP_ERR_INVALID; }
    // 1x1 convolutions are epilogue-bound (K is one or a few chunks): keep the accumulator at
    // <= 128 TMEM columns so that four CTAs share an SM and overlap each other's epilogues
    if (op.k == 1 && !persistent_tiling) while (n_tile > 128 && n_tile % 32 == 0) n_tile /= 2;
    // members of a persistent group share one ring of equal slots: every member's stage (A box + weight box)
    // stays <= 24 KB so that eight of them are in flight (the loads are L2-latency bound)
    if (persistent_tiling) {
        const int cap_n = op.cin == 64 ? 64 : 128;             // Cin >= 128: 16 KB A + 16 KB weights per stage, six stages
        while (n_tile > cap_n && n_tile % 32 == 0) n_tile /= 2;
    }
    // M tiles per CTA: 2 when that still leaves >= 2 CTAs per SM and TMEM stays <= 256 columns
    int m_tiles = 2, tn = 0, th = 0, tw = 0;
    {
        bool ok2 = pick_tile(Ho, Wo, 2, &tn, &th, &tw) && 2 * n_tile <= ((op.k == 1 && !persistent_tiling) ? 128 : 256) &&
                   !(persistent_tiling && op.cin != 32 && !(op.sm_share > 0.f && !for_group && env_int("HBP_PG_M2", 0)));
        // (persistent walkers want few, large work items: no minimum CTA count, no extra N split)
        if (ok2 && !persistent_tiling) {
            const long ctas = (long)((capP + tn - 1) / tn) * (Ho / th) * (Wo / tw) * (op.cout / n_tile);
            if (ctas < 2L * ctx->sm_count) ok2 = false;
        }
        if (!ok2) { m_tiles = 1; pick_tile(Ho, Wo, 1, &tn, &th, &tw); }
    }
    // too few CTAs: split N further (down to 32)
    {
        long ctas = (long)((capP + tn - 1) / tn) * (Ho / th) * (Wo / tw) * (op.cout / n_tile);
        while (!persistent_tiling && ctas < ctx->sm_count && n_tile % 32 == 0 && n_tile > 32) { n_tile /= 2; ctas *= 2; }
    }
    p.Ho = Ho; p.Wo = Wo; p.Cout = op.cout; p.up = op.up; p.relu = op.relu;
    p.tn = tn; p.th = th; p.tw = tw; p.m_tiles = m_tiles; p.n_tile = n_tile;
    p.chunk = op.cin == 32 ? 32 : 64;
    p.n_chunks = op.cin / p.chunk;
    p.ksz = op.k; p.stride = op.stride;
    p.tiles_w = Wo / tw; p.tiles_h = Ho / th;
    p.row_bytes = p.chunk * 2;
    p.a_stage_bytes = 128u * m_tiles * p.row_bytes;
    p.b_stage_bytes = ((uint32_t)n_tile * p.row_bytes + 1023u) & ~1023u;
    p.tx_bytes = p.a_stage_bytes + (uint32_t)n_tile * p.row_bytes;
    const uint32_t stage = p.a_stage_bytes + p.b_stage_bytes;
    const int k_iters = op.k * op.k * p.n_chunks;
    const uint32_t budget = stage * 4 <= 96 * 1024 ? 96 * 1024 : 200 * 1024;
    int stages = budget / stage;
    if (stages > 8) stages = 8;
    if (stages > k_iters) stages = k_iters;
    if (stages < 1) { delete pl; hbp_set_error("stage too large"); return HBP_ERR_INVALID; }
    p.stages = stages;
    uint32_t cols = 32;
    while (cols < (uint32_t)(m_tiles * n_tile)) cols *= 2;
    p.tmem_cols = cols;
    // instruction descriptor: D=F32 (bits 4-5 = 1), A=B=F16 (0), K-major both, N>>3 at 17, M>>4 at 24
    p.idesc = (1u << 4) | ((uint32_t)(n_tile >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
    p.bias = m.d_bias + op.b_off;
    p.res = op.res >= 0 ? m.bufs[m.tensors[op.res].buf] : nullptr;
    p.out = m.bufs[m.tensors[op.out].buf];
    pl->smem_bytes = (size_t)stages * stage + 16 * stages + 64 + (size_t)n_tile * 4 + 1024;
    pl->n_splits = op.cout / n_tile;

    const CUtensorMapSwizzle sw = p.row_bytes == 128 ? CU_TENSOR_MAP_SWIZZLE_128B : CU_TENSOR_MAP_SWIZZLE_64B;
    {
        cuuint64_t gdim[4] = {(cuuint64_t)ti.c, (cuuint64_t)ti.w, (cuuint64_t)ti.h, (cuuint64_t)capP};
        cuuint64_t gstr[3] = {(cuuint64_t)ti.c * 2, (cuuint64_t)ti.w * ti.c * 2, (cuuint64_t)ti.h * ti.w * ti.c * 2};
        const int s = op.stride;
        const int bw = (s == 2 && stride_mode() == 1) ? tw * s : tw;
        const int bh = (s == 2 && stride_mode() == 1) ? th * s : th;
        cuuint32_t box[4] = {(cuuint32_t)p.chunk, (cuuint32_t)bw, (cuuint32_t)bh, (cuuint32_t)tn};
        cuuint32_t est[4] = {1, (cuuint32_t)s, (cuuint32_t)s, 1};
        CUresult r = enc(&pl->tmA, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 4, m.bufs[ti.buf], gdim, gstr, box, est,
                         CU_TENSOR_MAP_INTERLEAVE_NONE, sw, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                         CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        if (r != CUDA_SUCCESS) {
            delete pl;
            hbp_set_error("cuTensorMapEncodeTiled(A) failed (%d) for %s box=(%d,%d,%d,%d) stride=%d", (int)r,
                          op.name.c_str(), p.chunk, bw, bh, tn, s);
            return HBP_ERR_CUDA;
        }
    }
    for (int q = 0; q < 4; ++q) pl->tmP.m[q] = pl->tmA;          // placeholders unless phase maps are in use
    if (op.stride == 2 && env_int("HBP_S2_PHASE", 1) && ti.h % 2 == 0 && ti.w % 2 == 0) {
        // Stride-2 taps as DENSE boxes.  Input pixel (2y+dy, 2x+dx) lies in the parity plane (dy&1, dx&1) of the
        // input at plane coordinates (y - (dy<0), x - (dx<0)); a parity plane is the tensor
        // (C, W/2, H/2, N) with strides (2C, 2WC, HWC) based at pixel (py, px).  A tiled map with
        // elementStrides = 2 walks the whole 2tw x 2th span of the box and keeps a quarter of it
        // (27 B/clk/SM measured on the stem's conv2); the plane view fetches only the rows it delivers.
        // Out-of-range plane coordinates (-1) are exactly the convolution's zero padding.
        for (int py = 0; py < 2; ++py)
            for (int px = 0; px < 2; ++px) {
                cuuint64_t gdim[4] = {(cuuint64_t)ti.c, (cuuint64_t)(ti.w / 2), (cuuint64_t)(ti.h / 2), (cuuint64_t)capP};
                cuuint64_t gstr[3] = {(cuuint64_t)ti.c * 4, (cuuint64_t)ti.w * ti.c * 4, (cuuint64_t)ti.h * ti.w * ti.c * 2};
                cuuint32_t box[4] = {(cuuint32_t)p.chunk, (cuuint32_t)tw, (cuuint32_t)th, (cuuint32_t)tn};
                cuuint32_t est[4] = {1, 1, 1, 1};
                CUresult r = enc(&pl->tmP.m[py * 2 + px], CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 4,
                                 m.bufs[ti.buf] + ((size_t)py * ti.w + px) * ti.c, gdim, gstr, box, est,
                                 CU_TENSOR_MAP_INTERLEAVE_NONE, sw, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                                 CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
                if (r != CUDA_SUCCESS) {
                    delete pl;
                    hbp_set_error("cuTensorMapEncodeTiled(A parity plane) failed (%d) for %s", (int)r, op.name.c_str());
                    return HBP_ERR_CUDA;
                }
            }
        p.phase_maps = 1;
    }
    {
        int st = encode_weights_map(enc, pl, m, op, p.chunk, n_tile, sw);
        if (st) { delete pl; return st; }
    }
    *out = pl;
    return HBP_OK;
}

struct UmmaGroup {
    UmmaGroup* next = nullptr;       // tables of the same launch for other batch sizes (a captured graph keeps pointing at its own)
    GroupEntry* d_table = nullptr;
    std::vector<GroupEntry> h;
    GroupHeader hdr;
    size_t smem = 0;
    int grid = 0;
    int P = -1;
    const void* tl = nullptr;
};

void umma_group_destroy(UmmaGroup* g) {
    while (g) {
        UmmaGroup* nx = g->next;
        if (g->d_table) cudaFree(g->d_table);
        delete g;
        g = nx;
    }
}

// one persistent launch over the work items of `members` (mode-0 plans in m.umma); `slot_index` keys the table
static int group_launch(hbp_ctx* ctx, HrnetModel& m, int slot_index, const int* members, int n, int P, cudaStream_t st) {
    if (n < 1 || n > kMaxGroup) { hbp_set_error("group of %d convolutions (max %d)", n, kMaxGroup); return HBP_ERR_INVALID; }
    if ((int)m.groups.size() < (int)m.ops.size()) m.groups.resize(m.ops.size(), nullptr);
    // one table per batch size: the graphs of other batch sizes keep replaying with theirs
    UmmaGroup* g = m.groups[slot_index];
    while (g && !(g->P == P && g->tl == (const void*)m.d_timeline)) g = g->next;
    if (!g) {
        g = new UmmaGroup();
        g->h.resize(n);
        HBP_CUDA(cudaMalloc(&g->d_table, sizeof(GroupEntry) * n));
        g->next = m.groups[slot_index];
        m.groups[slot_index] = g;
    }
    if (g->P != P || g->tl != (const void*)m.d_timeline) {
        // (re)build the table for this batch size.  Never inside a graph capture: the first forward of a
        // (batch, buffers) key runs eagerly, the capture of the second one finds the table in place.
        cudaStreamCaptureStatus cs = cudaStreamCaptureStatusNone;
        cudaStreamIsCapturing(st, &cs);
        if (cs != cudaStreamCaptureStatusNone) { hbp_set_error("group table rebuilt inside a capture"); return HBP_ERR_STATE; }
        int items = 0;
        uint32_t slot = 0, acc = 0;
        g->hdr.n = n;
        for (int i = 0; i < n; ++i) {
            const int oi = members[i];
            const UmmaPlan* pl = m.umma[oi];
            if (!pl || pl->prm.mode != 0) { hbp_set_error("group member %s has no mode-0 plan", m.ops[oi].name.c_str()); return HBP_ERR_STATE; }
            GroupEntry& e = g->h[i];
            e.tmA = pl->tmA; e.tmB = pl->tmB; e.tmP = pl->tmP; e.p = pl->prm;
            e.p.P = P;
            e.p.dbg_flags = env_int("HBP_PG_DBG", 0);
            e.p.tl = m.d_timeline ? m.d_timeline + 2 * oi : nullptr;
            const int tiles_n = (P + e.p.tn - 1) / e.p.tn;
            e.gx = tiles_n * e.p.tiles_h * e.p.tiles_w;
            e.gy = pl->n_splits;
            g->hdr.item_begin[i] = items;
            items += e.gx * e.gy;
            slot = std::max(slot, e.p.a_stage_bytes + e.p.b_stage_bytes);
            acc = std::max(acc, (uint32_t)(e.p.m_tiles * e.p.n_tile));
        }
        for (int i = n; i <= kMaxGroup; ++i) g->hdr.item_begin[i] = items;
        {
            int couts = 0;
            for (int i = 0; i < n; ++i) couts += (g->h[i].p.Cout + 3) & ~3;
            if (couts > kPGroupBias) { hbp_set_error("group has %d output channels (max %d)", couts, kPGroupBias); return HBP_ERR_INVALID; }
        }
        slot = (slot + 1023u) & ~1023u;
        int stages = (int)((uint32_t)env_int("HBP_PGROUP_SMEM_KB", 192) * 1024u / slot);
        if (stages > kMaxPStages) stages = kMaxPStages;
        if (stages < 2 || acc > 256) { hbp_set_error("group does not fit: slot %u B, %u accumulator columns", slot, acc); return HBP_ERR_INVALID; }
        uint32_t cols = 32;
        while (cols < 2 * acc) cols *= 2;
        g->hdr.stages = stages; g->hdr.slot_bytes = slot; g->hdr.acc_cols = acc; g->hdr.tmem_cols = cols;
        g->smem = (size_t)stages * slot + 1024;
        int sms = ctx->sm_count;
        if (n == 1 && m.ops[members[0]].sm_share > 0.f) sms = std::max(1, (int)(m.ops[members[0]].sm_share * ctx->sm_count + 0.5f));
        g->grid = std::min(items, sms);
        HBP_CUDA(cudaMemcpyAsync(g->d_table, g->h.data(), sizeof(GroupEntry) * n, cudaMemcpyHostToDevice, st));
        HBP_CUDA(cudaStreamSynchronize(st));
        g->P = P;
        g->tl = m.d_timeline;
        if (getenv("HBP_CONV_TRACE"))
            fprintf(stderr, "[pgroup] %s members=%d items=%d grid=%d stages=%d slot=%u acc_cols=%u tmem=%u smem=%zu\n",
                    m.ops[slot_index].name.c_str(), n, items, g->grid, stages, slot, acc, cols, g->smem);
    }
    static const int pdl = env_int("HBP_PDL", 1);
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3((unsigned)g->grid); cfg.blockDim = dim3(kPGroupThreads); cfg.dynamicSmemBytes = g->smem; cfg.stream = st;
    // high scheduling priority: a group with few, large CTAs (the last links of the stride-2 chains) must not
    // queue behind the thousands of small blocks of the upsample-add that runs beside it on another stream
    static int prio = 1;
    if (prio == 1) { int lo = 0, hi = 0; cudaDeviceGetStreamPriorityRange(&lo, &hi); prio = env_int("HBP_PGROUP_PRIO", 1) ? hi : 0; }
    cudaLaunchAttribute at[2];
    at[0].id = cudaLaunchAttributePriority;
    at[0].val.priority = prio;
    at[1].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    at[1].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = at; cfg.numAttrs = pdl ? 2 : 1;
    cudaLaunchKernelEx(&cfg, conv_umma_pgroup_kernel, (const GroupEntry*)g->d_table, g->hdr);
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return hbp_cuda_fail(e, "conv_umma_pgroup_kernel", __FILE__, __LINE__);
    return HBP_OK;
}

int umma_group_launch(hbp_ctx* ctx, HrnetModel& m, int group_index, int P, cudaStream_t st) {
    const HOp& gop = m.ops[group_index];
    return group_launch(ctx, m, group_index, gop.members.data(), (int)gop.members.size(), P, st);
}

static void launch_halo(dim3 grid, size_t smem, cudaStream_t st, const UmmaPlan* pl, const ConvParams& p) {
    static const int pdl = env_int("HBP_PDL", 1);
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = grid; cfg.blockDim = dim3(kHaloThreads); cfg.dynamicSmemBytes = smem; cfg.stream = st;
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    at[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = at; cfg.numAttrs = pdl ? 1 : 0;
    const int ks = p.chunk / 16;
#define HBP_HALO_CASE(K, M, S) if (ks == K && p.m_tiles == M && p.ksz == S) { cudaLaunchKernelEx(&cfg, conv_umma_halo_kernel<K, M, S>, pl->tmA, pl->tmB, pl->tmR, p); return; }
    HBP_HALO_CASE(2, 1, 3) HBP_HALO_CASE(2, 2, 3) HBP_HALO_CASE(4, 1, 3) HBP_HALO_CASE(4, 2, 3)
    HBP_HALO_CASE(2, 1, 1) HBP_HALO_CASE(2, 2, 1) HBP_HALO_CASE(4, 1, 1) HBP_HALO_CASE(4, 2, 1)
#undef HBP_HALO_CASE
}

int umma_launch(hbp_ctx* ctx, HrnetModel& m, int op_index, UmmaPlan* pl, int P, cudaStream_t st) {
    ConvParams p = pl->prm;
    p.P = P;
    p.tl = m.d_timeline ? m.d_timeline + 2 * op_index : nullptr;
    const int tiles_n = (P + p.tn - 1) / p.tn;
    dim3 grid((unsigned)(tiles_n * p.tiles_h * p.tiles_w), (unsigned)pl->n_splits);
    if (p.mode == 1) {
        // one persistent CTA per SM: every CTA walks ceil(n_tiles / ctas) or one fewer tiles
        p.n_tiles = (int)grid.x;
        int slots = pl->sm_budget / pl->n_splits;
        if (slots < 1) slots = 1;
        if ((int)grid.x > slots) {
            const int per = (p.n_tiles + slots - 1) / slots;
            grid.x = (unsigned)((p.n_tiles + per - 1) / per);
        }
    }
    static const bool trace = getenv("HBP_CONV_TRACE") != nullptr;
    cudaStreamCaptureStatus cap_status = cudaStreamCaptureStatusNone;
    if (trace) cudaStreamIsCapturing(st, &cap_status);
    if (trace && p.mode == 1 && cap_status == cudaStreamCaptureStatusNone) {
        {                               // every eager launch (the timed launches run inside a graph capture)
            long long* d = nullptr;
            cudaMalloc(&d, 64 * 256 * sizeof(long long));
            cudaMemset(d, 0, 64 * 256 * sizeof(long long));
            launch_halo(grid, pl->smem_bytes, st, pl, p);    // warm L2
            p.dbg = d;
            launch_halo(grid, pl->smem_bytes, st, pl, p);
            cudaStreamSynchronize(st);
            static long long h[64 * 256];
            cudaMemcpy(h, d, sizeof(h), cudaMemcpyDeviceToHost);
            cudaFree(d);
            const int n = grid.x < 64 ? (int)grid.x : 64;
            double acc[8] = {0};
            for (int i = 0; i < n; ++i) for (int k = 1; k < 8; ++k) acc[k] += (double)(h[i * 256 + k] - h[i * 256]);
            fprintf(stderr, "[trace] grid=(%u,%u) smem=%zu stages=%d m=%d n_tile=%d chunks=%d | cycles from CTA start: setup %.0f, A0 landed %.0f, "
                    "B0 landed %.0f, MMAs issued %.0f, accum ready %.0f, epilogue done %.0f, dealloc %.0f\n", grid.x, grid.y, pl->smem_bytes,
                    p.stages, p.m_tiles, p.n_tile, p.n_chunks, acc[1] / n, acc[2] / n, acc[3] / n, acc[4] / n, acc[5] / n, acc[6] / n, acc[7] / n);
            // timeline of CTA 0, cycles from its start: per tile j the MMA warp's (accumulator free, A landed, MMAs issued),
            // the epilogue's (accumulator ready, done) and when the producer requested the halo
            fprintf(stderr, "[timeline] tile: A requested | acc free, A landed, issued | epilogue ready, done\n");
            for (int j = 0; j < 16; ++j) {
                auto rel = [&](int k) { return h[k] ? (long long)(h[k] - h[0]) : -1LL; };
                if (!h[34 + 3 * j]) break;
                fprintf(stderr, "[timeline] %2d: %6lld | %6lld %6lld %6lld | %6lld %6lld | team: wait from %6lld, first ld %6lld, stores issued %6lld\n", j, rel(128 + j), rel(32 + 3 * j), rel(33 + 3 * j),
                        rel(34 + 3 * j), rel(96 + 2 * j), rel(97 + 2 * j), rel(160 + 4 * j), rel(161 + 4 * j), rel(162 + 4 * j));
            }
            p.dbg = nullptr;
            return HBP_OK;
        }
    }
    static const int persist0 = env_int("HBP_PERSIST0", 2);      // 1: only ops flagged `persist`; 2: also branch convs without a halo plan (256-channel 8x6 branch)
    if (p.mode == 1) launch_halo(grid, pl->smem_bytes, st, pl, p);
    else if (persist0 && (m.ops[op_index].persist || (persist0 > 1 && m.ops[op_index].sm_share > 0.f)) && p.m_tiles * p.n_tile <= 256) {
        // per-tap convolutions (stride 2, fused upsample) walk their tiles with persistent CTAs: a group of one
        const int member = op_index;
        return group_launch(ctx, m, op_index, &member, 1, P, st);
    } else {
        static const int pdl = env_int("HBP_PDL", 1);
        cudaLaunchConfig_t cfg = {};
        cfg.gridDim = grid; cfg.blockDim = dim3(kThreads); cfg.dynamicSmemBytes = pl->smem_bytes; cfg.stream = st;
        cudaLaunchAttribute at[1];
        at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
        at[0].val.programmaticStreamSerializationAllowed = 1;
        cfg.attrs = at; cfg.numAttrs = pdl ? 1 : 0;
        cudaLaunchKernelEx(&cfg, conv_umma_kernel, pl->tmA, pl->tmB, pl->tmP, p);
    }
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return hbp_cuda_fail(e, "conv_umma_kernel", __FILE__, __LINE__);
    return HBP_OK;
}

// ---------------------------------------------------------------------------
// chain launch (conv_umma_chain_kernel): host side
// ---------------------------------------------------------------------------
struct UmmaChain {
    UmmaChain* next = nullptr;           // per batch size, like UmmaGroup
    ChainEntry* d_table = nullptr;
    unsigned* d_flags = nullptr;         // [n_layers][n_tiles] + 2 words of launch state
    std::vector<ChainEntry> h;
    ConvParams prm;
    ChainHeader hdr;
    size_t smem = 0;
    int grid = 0, ksteps = 0;
    int P = -1;
    int unsupported = 0;                 // the members run as individual launches
    const void* tl = nullptr;
};

void umma_chain_destroy(UmmaChain* c) {
    while (c) {
        UmmaChain* nx = c->next;
        if (c->d_table) cudaFree(c->d_table);
        if (c->d_flags) cudaFree(c->d_flags);
        if (c->hdr.trace) cudaFree(c->hdr.trace);
        delete c;
        c = nx;
    }
}

static int encode_act_map(EncodeTiledFn enc, CUtensorMap* tm, const HrnetModel& m, const HTensor& t, int capP, int c_box, int w_box,
                          int h_box, int n_box, uint32_t row_bytes, const char* what) {
    cuuint64_t gdim[4] = {(cuuint64_t)t.c, (cuuint64_t)t.w, (cuuint64_t)t.h, (cuuint64_t)capP};
    cuuint64_t gstr[3] = {(cuuint64_t)t.c * 2, (cuuint64_t)t.w * t.c * 2, (cuuint64_t)t.h * t.w * t.c * 2};
    cuuint32_t box[4] = {(cuuint32_t)c_box, (cuuint32_t)w_box, (cuuint32_t)h_box, (cuuint32_t)n_box};
    cuuint32_t est[4] = {1, 1, 1, 1};
    CUresult r = enc(tm, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 4, m.bufs[t.buf], gdim, gstr, box, est, CU_TENSOR_MAP_INTERLEAVE_NONE,
                     row_bytes == 128 ? CU_TENSOR_MAP_SWIZZLE_128B : CU_TENSOR_MAP_SWIZZLE_64B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                     CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) { hbp_set_error("cuTensorMapEncodeTiled(%s) failed (%d)", what, (int)r); return HBP_ERR_CUDA; }
    return HBP_OK;
}

// returns HBP_OK after launching, 1 when the chain cannot run as one kernel (the caller issues the members one by one)
int umma_chain_launch(hbp_ctx* ctx, HrnetModel& m, int chain_index, int P, cudaStream_t st) {
    static const int enabled = env_int("HBP_CHAIN", 0);     // measured slower than the per-conv launches at batch 64 (profiles/r02_chain_kernel.md): off by default
    const HOp& cop = m.ops[chain_index];
    const int L = (int)cop.members.size();
    if (!enabled || L < 2 || L > kMaxChain) return 1;
    if ((int)m.chains.size() < (int)m.ops.size()) m.chains.resize(m.ops.size(), nullptr);
    UmmaChain* c = m.chains[chain_index];
    while (c && !c->unsupported && !(c->P == P && c->tl == (const void*)m.d_timeline)) c = c->next;
    if (!c) {
        c = new UmmaChain();
        c->hdr.trace = nullptr;
        c->next = m.chains[chain_index];
        m.chains[chain_index] = c;
    }
    if (c->unsupported) return 1;
    if (c->P != P || c->tl != (const void*)m.d_timeline) {
        cudaStreamCaptureStatus cs = cudaStreamCaptureStatusNone;
        cudaStreamIsCapturing(st, &cs);
        if (cs != cudaStreamCaptureStatusNone) { hbp_set_error("chain table rebuilt inside a capture"); return HBP_ERR_STATE; }
        // geometry from a member WITH a residual (its ring stage holds halo + residual tile): the most constrained plan
        int with_res = -1;
        for (int k = 0; k < L; ++k) if (m.ops[cop.members[k]].res >= 0) { with_res = k; break; }
        const HOp& op0 = m.ops[cop.members[with_res >= 0 ? with_res : 0]];
        for (int k = 0; k < L; ++k) {
            const HOp& o = m.ops[cop.members[k]];
            const HTensor &ti = m.tensors[o.in], &t0 = m.tensors[op0.in];
            if (o.k != 3 || o.stride != 1 || o.up != 1 || o.cin != op0.cin || o.cout != op0.cout || o.cin != o.cout || ti.h != t0.h || ti.w != t0.w ||
                !umma_supported(m, o)) { c->unsupported = 1; return 1; }
        }
        UmmaPlan pl;
        memset(&pl.prm, 0, sizeof(pl.prm));
        bool ok = false;
        int stt = plan_halo(ctx, m, op0, m.cap_P, &pl, &ok);
        if (stt) return stt;
        ConvParams& p = pl.prm;
        if (!ok || pl.n_splits != 1 || (with_res >= 0 && !p.res_smem) || p.bias_mma || p.res_mma) { c->unsupported = 1; return 1; }
        const int tiles_n = (P + p.tn - 1) / p.tn;
        const int n_tiles = tiles_n * p.tiles_h * p.tiles_w;
        int grid = std::min(pl.sm_budget, n_tiles);
        if (grid < 1) grid = 1;
        if (n_tiles / grid < 4) p.issuers = 1;             // both issuers need a last tile in every layer
        {
            const HTensor& to = m.tensors[op0.out];
            if ((n_tiles + grid - 1) / grid > kMaxChainTiles || (size_t)m.cap_P * to.h * to.w * to.c >= (size_t(1) << 32) ||
                tiles_n > 65535) { c->unsupported = 1; return 1; }
        }
        p.P = P; p.n_tiles = n_tiles; p.dbg = nullptr; p.tl = nullptr; p.dbg_flags = env_int("HBP_CHAIN_DBG", 0);
        c->prm = p;
        c->grid = grid;
        c->ksteps = p.chunk / 16;
        c->smem = pl.smem_bytes + (size_t)(L - 1) * p.n_tile * 4;
        p.b_sets = 1;
        if (p.b_resident) {
            // a second resident weight set (the next layer's, prefetched a whole layer ahead) when it fits
            const size_t b_all = (size_t)p.b_slots * 3 * p.b_stage_bytes;
            if (env_int("HBP_CHAIN_BSETS", 2) >= 2 && c->smem + b_all <= 212 * 1024 && 2 * p.b_slots <= kMaxBSlots) {
                p.b_sets = 2;
                c->smem += b_all;
                p.b_slots *= 2;
            }
        }
        c->prm = p;
        if (c->smem > 212 * 1024) { c->unsupported = 1; return 1; }     // (the kernel's tables take 12 KB of static shared memory)
        c->h.assign(L, ChainEntry());
        EncodeTiledFn enc = get_encode();
        const int hw = kHaloW;
        std::vector<int> out_tensor(L);
        for (int k = 0; k < L; ++k) {
            const HOp& o = m.ops[cop.members[k]];
            ChainEntry& e = c->h[k];
            int s1 = encode_act_map(enc, &e.tmA, m, m.tensors[o.in], m.cap_P, p.chunk, hw, p.rs, p.tn, p.row_bytes, "chain A");
            if (s1) return s1;
            e.tmR = e.tmA;
            if (o.res >= 0) {
                s1 = encode_act_map(enc, &e.tmR, m, m.tensors[o.res], m.cap_P, (int)(p.r_row_bytes / 2), 8, p.th, p.tn, p.r_row_bytes, "chain residual");
                if (s1) return s1;
            }
            UmmaPlan tmp;
            s1 = encode_weights_map(enc, &tmp, m, o, p.chunk, p.n_tile, p.row_bytes == 128 ? CU_TENSOR_MAP_SWIZZLE_128B : CU_TENSOR_MAP_SWIZZLE_64B);
            if (s1) return s1;
            e.tmB = tmp.tmB;
            e.bias = m.d_bias + o.b_off;
            e.out = m.bufs[m.tensors[o.out].buf];
            e.relu = o.relu; e.has_res = o.res >= 0 ? 1 : 0;
            e.res_layer = -1;
            for (int j = 0; j < k; ++j) if (o.res >= 0 && out_tensor[j] == o.res) e.res_layer = j;
            // the input of layer k must be the output of layer k-1 (that is what the flags order)
            if (k > 0 && o.in != out_tensor[k - 1]) { c->unsupported = 1; return 1; }
            out_tensor[k] = o.out;
            e.tl = m.d_timeline ? m.d_timeline + 2 * cop.members[k] : nullptr;
        }
        if (!c->d_table) HBP_CUDA(cudaMalloc(&c->d_table, sizeof(ChainEntry) * kMaxChain));
        if (c->d_flags) { HBP_CUDA(cudaStreamSynchronize(st)); cudaFree(c->d_flags); c->d_flags = nullptr; }
        const size_t n_flags = (size_t)L * n_tiles + 2;
        HBP_CUDA(cudaMalloc(&c->d_flags, n_flags * sizeof(unsigned)));
        HBP_CUDA(cudaMemsetAsync(c->d_flags, 0, n_flags * sizeof(unsigned), st));
        HBP_CUDA(cudaMemcpyAsync(c->d_table, c->h.data(), sizeof(ChainEntry) * L, cudaMemcpyHostToDevice, st));
        HBP_CUDA(cudaStreamSynchronize(st));
        c->hdr.n_layers = L;
        c->hdr.flags = c->d_flags;
        c->hdr.state = c->d_flags + (size_t)L * n_tiles;
        c->hdr.trace = nullptr;
        if (getenv("HBP_CHAIN_TRACE") && cop.name.find(getenv("HBP_CHAIN_TRACE")) != std::string::npos) {
            HBP_CUDA(cudaMalloc(&c->hdr.trace, (size_t)grid * L * 2 * sizeof(unsigned long long)));
            HBP_CUDA(cudaMemset(c->hdr.trace, 0, (size_t)grid * L * 2 * sizeof(unsigned long long)));
        }
        c->P = P;
        c->tl = m.d_timeline;
        if (getenv("HBP_CONV_TRACE"))
            fprintf(stderr, "[chain] %s layers=%d tiles=%d grid=%d (%d..%d tiles per CTA) m=%d n_tile=%d a_stages=%d b_slots=%d resident=%d issuers=%d teams=%d smem=%zu\n",
                    cop.name.c_str(), L, n_tiles, grid, n_tiles / grid, (n_tiles + grid - 1) / grid, p.m_tiles, p.n_tile, p.a_stages, p.b_slots,
                    p.b_resident, p.issuers, p.teams, c->smem);
    }
    static const int pdl = env_int("HBP_PDL", 1);
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3((unsigned)c->grid); cfg.blockDim = dim3(kHaloThreads); cfg.dynamicSmemBytes = c->smem; cfg.stream = st;
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    at[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = at; cfg.numAttrs = pdl ? 1 : 0;
    const int ks = c->ksteps, mt = c->prm.m_tiles;
    if (ks == 2 && mt == 1) cudaLaunchKernelEx(&cfg, conv_umma_chain_kernel<2, 1>, (const ChainEntry*)c->d_table, c->prm, c->hdr);
    else if (ks == 2 && mt == 2) cudaLaunchKernelEx(&cfg, conv_umma_chain_kernel<2, 2>, (const ChainEntry*)c->d_table, c->prm, c->hdr);
    else if (ks == 4 && mt == 1) cudaLaunchKernelEx(&cfg, conv_umma_chain_kernel<4, 1>, (const ChainEntry*)c->d_table, c->prm, c->hdr);
    else if (ks == 4 && mt == 2) cudaLaunchKernelEx(&cfg, conv_umma_chain_kernel<4, 2>, (const ChainEntry*)c->d_table, c->prm, c->hdr);
    else { c->unsupported = 1; return 1; }
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return hbp_cuda_fail(e, "conv_umma_chain_kernel", __FILE__, __LINE__);
    if (c->hdr.trace) {
        cudaStreamCaptureStatus cs = cudaStreamCaptureStatusNone;
        cudaStreamIsCapturing(st, &cs);
        static int dumps = 0;
        if (cs == cudaStreamCaptureStatusNone && dumps < 4) {
            // eager launches only: per CTA the start of its first and the end of its last tile of every layer, us from the launch's first tile
            ++dumps;
            cudaDeviceSynchronize();
            std::vector<unsigned long long> h((size_t)c->grid * L * 2);
            cudaMemcpy(h.data(), c->hdr.trace, h.size() * sizeof(unsigned long long), cudaMemcpyDeviceToHost);
            unsigned long long t0 = ~0ull;
            for (int i = 0; i < c->grid; ++i) if (h[(size_t)i * L * 2] && h[(size_t)i * L * 2] < t0) t0 = h[(size_t)i * L * 2];
            fprintf(stderr, "[chaintrace] %s grid=%d layers=%d: per CTA 'start0 | end of layer 0..%d' (us)\n", cop.name.c_str(), c->grid, L, L - 1);
            for (int i = 0; i < c->grid; ++i) {
                fprintf(stderr, "[chaintrace] cta %3d tiles %3d: %7.1f |", i, (int)(((long long)(i + 1) * c->prm.n_tiles) / c->grid - ((long long)i * c->prm.n_tiles) / c->grid),
                        (double)(h[(size_t)i * L * 2] - t0) * 1e-3);
                for (int l = 0; l < L; ++l) fprintf(stderr, " %7.1f", (double)(h[((size_t)i * L + l) * 2 + 1] - t0) * 1e-3);
                fprintf(stderr, "\n");
            }
        }
    }
    return HBP_OK;
}
