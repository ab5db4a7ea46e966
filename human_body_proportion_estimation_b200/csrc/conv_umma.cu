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
    int acc_bufs;                // halo mode: 1 or 2 accumulator buffers in TMEM
    long long* dbg;              // optional phase timestamps (8 per CTA, first 64 CTAs), bring-up only
};

struct UmmaPlan {
    CUtensorMap tmA, tmB;
    ConvParams prm;
    size_t smem_bytes;
    int n_splits;
    int occ;                     // resident CTAs per SM (halo mode: sizes the persistent grid)
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
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
    // try_wait suspends in hardware for a bounded time; a pipeline that has not
    // advanced for ~2 s is a bug (wrong expect_tx byte count, bad tensor map):
    // trap instead of hanging the GPU.
    long long t0 = 0;
    for (uint32_t spins = 0;; ++spins) {
        uint32_t done;
        asm volatile(
            "{\n"
            ".reg .pred p;\n"
            "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n"
            "selp.u32 %0, 1, 0, p;\n"
            "}\n" : "=r"(done) : "r"(bar), "r"(parity) : "memory");
        if (done) return;
        if ((spins & 1023u) == 1023u) {
            const long long now = clock64();
            if (t0 == 0) t0 = now;
            else if (now - t0 > 4000000000LL) __trap();
        }
    }
}
__device__ __forceinline__ void tma_load_4d(uint32_t dst, const CUtensorMap* map, uint32_t bar, int c0, int c1, int c2, int c3) {
    asm volatile(
        "cp.async.bulk.tensor.4d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2];"
        ::"r"(dst), "l"(map), "r"(bar), "r"(c0), "r"(c1), "r"(c2), "r"(c3) : "memory");
}
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

constexpr int kThreads = 192;    // warp 0 TMA, warp 1 MMA + TMEM alloc, warps 2-5 epilogue

__global__ void __launch_bounds__(kThreads)
conv_umma_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
                 const ConvParams p) {
    extern __shared__ __align__(1024) uint8_t smem_raw[];
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
    int t = blockIdx.x;
    const int tile_w = t % p.tiles_w; t /= p.tiles_w;
    const int tile_h = t % p.tiles_h; t /= p.tiles_h;
    const int n0 = t * p.tn, h0 = tile_h * p.th, w0 = tile_w * p.tw;
    const int n_off = blockIdx.y * p.n_tile;

    if (warp == 0 && lane == 0) {
        asm volatile("prefetch.tensormap [%0];" ::"l"(&tmA) : "memory");
        asm volatile("prefetch.tensormap [%0];" ::"l"(&tmB) : "memory");
        for (int s = 0; s < p.stages; ++s) {
            mbar_init(full_bar + 8u * s, 1);
            mbar_init(empty_bar + 8u * s, 1);
        }
        mbar_init(tmem_full_bar, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 1) tmem_alloc(tmem_slot, p.tmem_cols);
    for (int i = threadIdx.x; i < p.n_tile; i += kThreads) s_bias[i] = p.bias[blockIdx.y * p.n_tile + i];
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    uint32_t tmem_base;
    asm volatile("ld.shared.b32 %0, [%1];" : "=r"(tmem_base) : "r"(tmem_slot));
    tmem_base = __shfl_sync(0xffffffffu, tmem_base, 0);

    if (warp == 0) {
        // ===== TMA producer (warp-wide control flow, one elected lane issues) =====
        const int pad = p.ksz / 2;
        int s = 0;
        uint32_t ph = 0;
        for (int tap = 0; tap < taps; ++tap) {
            const int dy = tap / p.ksz - pad, dx = tap % p.ksz - pad;
            for (int cc = 0; cc < p.n_chunks; ++cc) {
                mbar_wait(empty_bar + 8u * s, ph ^ 1u);
                if (elect_one()) {
                    mbar_expect_tx(full_bar + 8u * s, p.tx_bytes);
                    tma_load_4d(a_base + s * p.a_stage_bytes, &tmA, full_bar + 8u * s,
                                cc * p.chunk, w0 * p.stride + dx, h0 * p.stride + dy, n0);
                    tma_load_2d(b_base + s * p.b_stage_bytes, &tmB, full_bar + 8u * s,
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
                            *reinterpret_cast<uint4*>(p.out + o) = *reinterpret_cast<const uint4*>(&pk[0]);
                            *reinterpret_cast<uint4*>(p.out + o + 8) = *reinterpret_cast<const uint4*>(&pk[4]);
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
}


constexpr int kHaloW = 10;       // 8 output columns + 1 halo column each side
constexpr int kMaxChunks = 8;

// 3x3 stride-1 convolution, halo mode, persistent CTAs.  Tile = tn images x th rows x 8
// columns.  In shared memory a chunk is the TMA box (chunk channels, 10, rs = th+2, tn): pixel
// rows of `row_bytes` ordered [n][h][w], i.e. "stacked" image rows q = n*rs + h of 10 pixels
// each.  MMA row r of M-tile mt is pixel column r%8 of stacked row g = mt*16 + r/8; tap (dy,dx)
// reads stacked row g+dy, column r%8+dx: a K-major operand with start offset
// ((mt*16+dy)*10+dx)*row_bytes and an 8-row-group stride of 10 rows.  Stacked rows that are
// halo rows (g % rs >= th) yield garbage accumulator rows which the epilogue skips.
//
// Each CTA walks tiles blockIdx.x, +gridDim.x, ... (grid = SMs x resident CTAs, so there is
// no partial last wave).  The accumulator is double-buffered in TMEM when it fits
// (p.acc_bufs == 2): the MMA lane starts tile j+1 while the epilogue warps drain tile j; the
// producer refills the halo buffers as soon as tile j's MMAs have retired.  The residual and
// the bias are fetched before the epilogue waits for the accumulator.
constexpr int kResPrefetch = 8;   // uint4 (8 halfs) of residual a thread may hold in flight

__global__ void __launch_bounds__(kThreads)
conv_umma_halo_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
                      const ConvParams p) {
    extern __shared__ __align__(1024) uint8_t smem_raw[];
    const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
    const uint32_t a_base = smem_base;
    const uint32_t b_base = a_base + p.n_chunks * p.a_chunk_bytes;
    const uint32_t bar_base = b_base + p.stages * p.b_stage_bytes;
    const uint32_t a_full = bar_base;                                   // kMaxChunks x 8 B
    const uint32_t b_full = a_full + 8u * kMaxChunks;
    const uint32_t b_empty = b_full + 8u * p.stages;
    const uint32_t acc_full = b_empty + 8u * p.stages;                  // 2 x 8 B
    const uint32_t acc_empty = acc_full + 16u;                          // 2 x 8 B
    const uint32_t tmem_slot = acc_empty + 16u;
    float* s_bias = reinterpret_cast<float*>(smem_raw + (tmem_slot + 16u - smem_u32(smem_raw)));

    const int warp = uniform_warp_idx(), lane = threadIdx.x & 31;
    const int n_off = blockIdx.y * p.n_tile;
    const int tiles_per_img = p.tiles_w * p.tiles_h;
    long long* dbg = (p.dbg && blockIdx.x < 64 && blockIdx.y == 0) ? p.dbg + blockIdx.x * 32 : nullptr;
    if (dbg && threadIdx.x == 0) dbg[0] = clock64();

    if (warp == 0 && lane == 0) {
        asm volatile("prefetch.tensormap [%0];" ::"l"(&tmA) : "memory");
        asm volatile("prefetch.tensormap [%0];" ::"l"(&tmB) : "memory");
        for (int c = 0; c < p.n_chunks; ++c) mbar_init(a_full + 8u * c, 1);
        for (int s = 0; s < p.stages; ++s) { mbar_init(b_full + 8u * s, 1); mbar_init(b_empty + 8u * s, 1); }
        for (int i = 0; i < 2; ++i) { mbar_init(acc_full + 8u * i, 1); mbar_init(acc_empty + 8u * i, 128); }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 1) tmem_alloc(tmem_slot, p.tmem_cols);
    for (int i = threadIdx.x; i < p.n_tile; i += kThreads) s_bias[i] = p.bias[n_off + i];
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    uint32_t tmem_base;
    asm volatile("ld.shared.b32 %0, [%1];" : "=r"(tmem_base) : "r"(tmem_slot));
    tmem_base = __shfl_sync(0xffffffffu, tmem_base, 0);
    if (dbg && threadIdx.x == 0) dbg[1] = clock64();
    const uint32_t acc_stride = (uint32_t)(p.m_tiles * p.n_tile);      // TMEM columns per accumulator buffer

    if (warp == 0) {
        // ===== TMA producer (warp-wide control flow, one elected lane issues) =====
        int s = 0, j = 0;
        uint32_t ph = 0;
        for (int tile = blockIdx.x; tile < p.n_tiles; tile += gridDim.x, ++j) {
            const int tile_w = tile % p.tiles_w, tile_h = (tile / p.tiles_w) % p.tiles_h;
            const int n0 = (tile / tiles_per_img) * p.tn, h0 = tile_h * p.th, w0 = tile_w * 8;
            // the halo buffers are free once the previous tile's MMAs have retired
            if (j > 0) mbar_wait(acc_full + 8u * ((j - 1) % p.acc_bufs), (uint32_t)(((j - 1) / p.acc_bufs) & 1));
            if (elect_one()) {
                for (int cc = 0; cc < p.n_chunks; ++cc) {
                    mbar_expect_tx(a_full + 8u * cc, p.a_box_bytes);
                    tma_load_4d(a_base + cc * p.a_chunk_bytes, &tmA, a_full + 8u * cc, cc * p.chunk, w0 - 1, h0 - 1, n0);
                }
            }
            __syncwarp();
            for (int cc = 0; cc < p.n_chunks; ++cc)
                for (int tap = 0; tap < 9; ++tap) {
                    mbar_wait(b_empty + 8u * s, ph ^ 1u);
                    if (elect_one()) {
                        mbar_expect_tx(b_full + 8u * s, (uint32_t)p.n_tile * p.row_bytes);
                        tma_load_2d(b_base + s * p.b_stage_bytes, &tmB, b_full + 8u * s, cc * p.chunk, tap * p.Cout + n_off);
                    }
                    __syncwarp();
                    if (++s == p.stages) { s = 0; ph ^= 1u; }
                }
        }
    } else if (warp == 1) {
        // ===== MMA issuer (warp-wide control flow, one elected lane issues) =====
        {
            const int ksteps = p.chunk / 16;
            const uint64_t da0 = make_desc(0, p.row_bytes, kHaloW * p.row_bytes);   // A: 8-row groups 10 rows apart
            const uint64_t db0 = make_desc(0, p.row_bytes);
            const uint32_t row16 = p.row_bytes >> 4;                                // one pixel row in 16-byte units
            const uint32_t mt_step = 16u * kHaloW * row16;
            int s = 0, j = 0;
            uint32_t ph = 0;
            for (int tile = blockIdx.x; tile < p.n_tiles; tile += gridDim.x, ++j) {
                const int ab = j % p.acc_bufs;
                // accumulator buffer `ab` must have been drained by the epilogue of tile j - acc_bufs
                if (j >= p.acc_bufs) {
                    mbar_wait(acc_empty + 8u * ab, (uint32_t)(((j - p.acc_bufs) / p.acc_bufs) & 1));
                    tc_fence_after();
                }
                const uint32_t d_base = tmem_base + ab * acc_stride;
                uint32_t first = 0;
                for (int cc = 0; cc < p.n_chunks; ++cc) {
                    mbar_wait(a_full + 8u * cc, (uint32_t)(j & 1));
                    if (dbg && cc == 0 && j == 0 && lane == 0) dbg[2] = clock64();
                    const uint64_t a_c = da0 + ((a_base + cc * p.a_chunk_bytes) >> 4);
                    for (int dy = 0; dy < 3; ++dy)
                        for (int dx = 0; dx < 3; ++dx) {
                            mbar_wait(b_full + 8u * s, ph);
                            if (dbg && !first && j == 0 && lane == 0) dbg[3] = clock64();
                            if (dbg && j == 0 && cc == 0 && lane == 0) dbg[8 + dy * 3 + dx] = clock64();
                            tc_fence_after();
                            const uint64_t bd0 = db0 + ((b_base + s * p.b_stage_bytes) >> 4);
                            const uint64_t a_t = a_c + (uint32_t)(dy * kHaloW + dx) * row16;
                            if (elect_one()) {
                                for (int mt = 0; mt < p.m_tiles; ++mt) {
                                    const uint64_t ad = a_t + mt * mt_step;
                                    const uint32_t dt = d_base + mt * p.n_tile;
                                    umma_f16(dt, ad, bd0, p.idesc, first);
                                    umma_f16(dt, ad + 2, bd0 + 2, p.idesc, 1u);
                                    if (ksteps == 4) {
                                        umma_f16(dt, ad + 4, bd0 + 4, p.idesc, 1u);
                                        umma_f16(dt, ad + 6, bd0 + 6, p.idesc, 1u);
                                    }
                                }
                                umma_commit(b_empty + 8u * s);
                            }
                            __syncwarp();
                            first = 1u;
                            if (dbg && j == 0 && cc == 0 && lane == 0) dbg[20 + dy * 3 + dx] = clock64();
                            if (++s == p.stages) { s = 0; ph ^= 1u; }
                        }
                }
                if (elect_one()) umma_commit(acc_full + 8u * ab);
                __syncwarp();
                if (dbg && j == 0 && lane == 0) dbg[4] = clock64();
            }
        }
    } else {
        // ===== epilogue: TMEM -> registers -> (+bias, +residual, ReLU) -> global =====
        const int grp = warp & 3;
        const int r = grp * 32 + lane;
        const int chunks8 = p.n_tile >> 3;                      // uint4 per pixel row
        const bool prefetch = p.res != nullptr && p.m_tiles * chunks8 <= kResPrefetch;
        int j = 0;
        for (int tile = blockIdx.x; tile < p.n_tiles; tile += gridDim.x, ++j) {
            const int tile_w = tile % p.tiles_w, tile_h = (tile / p.tiles_w) % p.tiles_h;
            const int n0 = (tile / tiles_per_img) * p.tn, h0 = tile_h * p.th, w0 = tile_w * 8;
            const int ab = j % p.acc_bufs;
            bool valid[2];
            size_t obase[2];
#pragma unroll
            for (int mt = 0; mt < 2; ++mt) {
                const int g = mt * 16 + (r >> 3), col = r & 7;
                const int nn = g / p.rs, hh = g - nn * p.rs;
                const int n = n0 + nn, ho = h0 + hh, wo = w0 + col;
                valid[mt] = mt < p.m_tiles && nn < p.tn && hh < p.th && n < p.P && wo < p.Wo;
                obase[mt] = ((((size_t)n * p.Ho + ho) * p.Wo) + wo) * p.Cout + n_off;
            }
            uint4 rq[kResPrefetch];
            if (prefetch) {
#pragma unroll
                for (int q = 0; q < kResPrefetch; ++q) {
                    const int mt = q / chunks8, c8 = q - mt * chunks8;
                    if (mt < p.m_tiles && valid[mt & 1])
                        rq[q] = *reinterpret_cast<const uint4*>(p.res + obase[mt & 1] + c8 * 8);
                }
            }
            mbar_wait(acc_full + 8u * ab, (uint32_t)((j / p.acc_bufs) & 1));
            if (dbg && warp == 2 && lane == 0 && j == 0) dbg[5] = clock64();
            tc_fence_after();
            for (int mt = 0; mt < p.m_tiles; ++mt) {
                for (int c0 = 0; c0 < p.n_tile; c0 += 16) {
                    uint32_t rr[16];
                    tmem_ld16(tmem_base + ((uint32_t)(grp * 32) << 16) + ab * acc_stride + (uint32_t)(mt * p.n_tile + c0), rr);
                    tmem_ld_wait();
                    if (!valid[mt]) continue;
                    float x[16];
#pragma unroll
                    for (int q = 0; q < 16; ++q) x[q] = __uint_as_float(rr[q]) + s_bias[c0 + q];
                    const size_t o = obase[mt] + c0;
                    if (p.res) {
                        uint4 q0, q1;
                        if (prefetch) {
                            const int qi = mt * chunks8 + (c0 >> 3);
                            // constant-index selects keep rq[] in registers
                            q0 = rq[0]; q1 = rq[1];
#pragma unroll
                            for (int t = 0; t < kResPrefetch; t += 2)
                                if (t == qi) { q0 = rq[t]; q1 = rq[t + 1]; }
                        } else {
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
                    *reinterpret_cast<uint4*>(p.out + o) = *reinterpret_cast<const uint4*>(&pk[0]);
                    *reinterpret_cast<uint4*>(p.out + o + 8) = *reinterpret_cast<const uint4*>(&pk[4]);
                }
            }
            // this thread's TMEM reads of buffer `ab` are complete: hand it back to the MMA lane
            tc_fence_before();
            asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(acc_empty + 8u * ab) : "memory");
            if (dbg && warp == 2 && lane == 0 && j == 0) dbg[6] = clock64();
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 1) {
        tc_fence_after();
        tmem_dealloc(tmem_base, p.tmem_cols);
        if (dbg && lane == 0) dbg[7] = clock64();
    }
}

// halo tile shape for an Ho x Wo map: tn images x th rows (x 8 columns), m_tiles M-tiles.
// Returns the fraction of MMA rows that are real output pixels (rows only; columns add Wo/(8*ceil(Wo/8))).
double pick_halo_tile(int Ho, int* tn, int* th, int* m_tiles) {
    double best = 0;
    for (int m = 2; m >= 1; --m)                                // ties go to two M-tiles (weights shared)
        for (int n = 1; n <= 6; ++n)
            for (int h = Ho < 16 * m ? Ho : 16 * m; h >= 1; --h) {
                if (Ho % h) continue;
                if (n > 1 && h != Ho) continue;                 // stacked images need whole images
                const int rs = h + 2;
                // all groups of the tile must be covered: 16m >= n*rs - 2, and the reads stay in the buffer
                if (16 * m < n * rs - 2) continue;
                const double frac = (double)(n * h) / (16.0 * m);
                if (frac > best + 1e-9) { best = frac; *tn = n; *th = h; *m_tiles = m; }
                break;                                          // largest h for this (m, n)
            }
    return best;
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

// halo-mode plan (3x3, stride 1).  Returns HBP_OK with *ok = false when the shape does not fit.
static int plan_halo(hbp_ctx* ctx, HrnetModel& m, const HOp& op, int capP, UmmaPlan* pl, bool* ok) {
    *ok = false;
    const HTensor& ti = m.tensors[op.in];
    const int Ho = ti.h, Wo = ti.w;
    ConvParams& p = pl->prm;
    int tn = 1, th = 0, m_tiles = 1;
    const double frac = pick_halo_tile(Ho, &tn, &th, &m_tiles);
    if (frac < 0.5) return HBP_OK;
    const int chunk = op.cin == 32 ? 32 : 64, n_chunks = op.cin / chunk;
    if (n_chunks > kMaxChunks) return HBP_OK;
    const uint32_t row_bytes = chunk * 2;
    const int tiles_w = (Wo + 7) / 8, tiles_h = Ho / th;
    int n_tile = 0;
    for (int c = op.cout < 256 ? op.cout : 256; c >= 16; c -= 16)
        if (op.cout % c == 0 && m_tiles * c <= 512) { n_tile = c; break; }
    if (!n_tile) return HBP_OK;
    // prefer one M-tile when two would leave SMs idle
    long tiles = (long)((capP + tn - 1) / tn) * tiles_h * tiles_w;
    if (m_tiles == 2 && tn == 1 && th == 32 && tiles * (op.cout / n_tile) < 2L * ctx->sm_count) { th = 16; m_tiles = 1; }
    // tiles are scarce (low-resolution branches at small batch): latency per CTA matters more than
    // the share of useful MMA rows -- take the smallest tile (one image, one M-tile)
    if (tiles * (op.cout / n_tile) < ctx->sm_count && (tn > 1 || m_tiles > 1)) {
        tn = 1; m_tiles = 1;
        th = Ho < 16 ? Ho : 16;
        while (Ho % th) --th;
    }
    const int tiles_h2 = Ho / th;
    tiles = (long)((capP + tn - 1) / tn) * tiles_h2 * tiles_w;
    long ctas = tiles * (op.cout / n_tile);
    while (ctas < ctx->sm_count && n_tile % 32 == 0 && n_tile > 32) { n_tile /= 2; ctas *= 2; }
    const int rs = th + 2;
    const uint32_t a_box_bytes = (uint32_t)(kHaloW * rs * tn) * row_bytes;
    uint32_t a_chunk_bytes = (uint32_t)((16 * m_tiles + 2) * kHaloW) * row_bytes;
    if (a_chunk_bytes < a_box_bytes) a_chunk_bytes = a_box_bytes;
    a_chunk_bytes = (a_chunk_bytes + 1023u) & ~1023u;
    const uint32_t b_stage = ((uint32_t)n_tile * row_bytes + 1023u) & ~1023u;
    const uint32_t a_total = a_chunk_bytes * n_chunks;
    const int k_iters = 9 * n_chunks;
    int stages = 8;
    if (stages > k_iters) stages = k_iters;
    while (stages > 2 && a_total + stages * b_stage > 200 * 1024) --stages;
    if (a_total + stages * b_stage > 200 * 1024) return HBP_OK;
    // small CTAs: keep shared memory low enough for >= 3 CTAs per SM
    while (stages > 4 && a_total + stages * b_stage > 72 * 1024) --stages;
    if (kHaloW > 256 || rs > 256 || tn > 256) return HBP_OK;

    p.Ho = Ho; p.Wo = Wo; p.Cout = op.cout; p.up = 1; p.relu = op.relu;
    p.tn = tn; p.th = th; p.tw = 8; p.m_tiles = m_tiles; p.n_tile = n_tile;
    p.chunk = chunk; p.n_chunks = n_chunks; p.ksz = 3; p.stride = 1;
    p.tiles_w = tiles_w; p.tiles_h = tiles_h2;
    p.row_bytes = row_bytes;
    p.a_stage_bytes = 0; p.b_stage_bytes = b_stage; p.tx_bytes = 0;
    p.stages = stages;
    p.acc_bufs = 2 * m_tiles * n_tile <= 256 ? 2 : 1;
    uint32_t cols = 32;
    while (cols < (uint32_t)(p.acc_bufs * m_tiles * n_tile)) cols *= 2;
    p.tmem_cols = cols;
    p.idesc = (1u << 4) | ((uint32_t)(n_tile >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
    p.mode = 1; p.rs = rs; p.a_chunk_bytes = a_chunk_bytes; p.a_box_bytes = a_box_bytes;
    p.bias = m.d_bias + op.b_off;
    p.res = op.res >= 0 ? m.bufs[m.tensors[op.res].buf] : nullptr;
    p.out = m.bufs[m.tensors[op.out].buf];
    pl->smem_bytes = (size_t)a_total + (size_t)stages * b_stage + 8 * kMaxChunks + 16 * stages + 64 + (size_t)n_tile * 4 + 1024;
    pl->n_splits = op.cout / n_tile;
    {
        // resident CTAs per SM: registers, shared memory (+1 KB the runtime reserves per CTA),
        // threads and TMEM columns
        cudaFuncAttributes fa;
        int regs = 80;
        if (cudaFuncGetAttributes(&fa, conv_umma_halo_kernel) == cudaSuccess && fa.numRegs > 0) regs = fa.numRegs;
        const int regs_per_cta = ((regs + 7) / 8 * 8) * kThreads;
        int occ = 65536 / regs_per_cta;
        const int by_smem = (int)((227 * 1024) / (pl->smem_bytes + 1024));
        const int by_threads = 2048 / kThreads;
        const int by_tmem = 512 / (int)cols;
        if (occ > by_smem) occ = by_smem;
        if (occ > by_threads) occ = by_threads;
        if (occ > by_tmem) occ = by_tmem;
        pl->occ = occ < 1 ? 1 : occ;
        if (getenv("HBP_CONV_TRACE"))
            fprintf(stderr, "[plan] %s halo tile tn=%d th=%d m=%d n_tile=%d stages=%d acc_bufs=%d tmem=%u regs=%d smem=%zu occ=%d\n",
                    op.name.c_str(), tn, th, m_tiles, n_tile, stages, p.acc_bufs, cols, regs, pl->smem_bytes, pl->occ);
    }

    EncodeTiledFn enc = get_encode();
    const CUtensorMapSwizzle sw = row_bytes == 128 ? CU_TENSOR_MAP_SWIZZLE_128B : CU_TENSOR_MAP_SWIZZLE_64B;
    cuuint64_t gdim[4] = {(cuuint64_t)ti.c, (cuuint64_t)ti.w, (cuuint64_t)ti.h, (cuuint64_t)capP};
    cuuint64_t gstr[3] = {(cuuint64_t)ti.c * 2, (cuuint64_t)ti.w * ti.c * 2, (cuuint64_t)ti.h * ti.w * ti.c * 2};
    cuuint32_t box[4] = {(cuuint32_t)chunk, (cuuint32_t)kHaloW, (cuuint32_t)rs, (cuuint32_t)tn};
    cuuint32_t est[4] = {1, 1, 1, 1};
    CUresult r = enc(&pl->tmA, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 4, m.bufs[ti.buf], gdim, gstr, box, est,
                     CU_TENSOR_MAP_INTERLEAVE_NONE, sw, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                     CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) {
        hbp_set_error("cuTensorMapEncodeTiled(A halo) failed (%d) for %s box=(%d,%d,%d,%d)", (int)r, op.name.c_str(),
                      chunk, kHaloW, rs, tn);
        return HBP_ERR_CUDA;
    }
    int st = encode_weights_map(enc, pl, m, op, chunk, n_tile, sw);
    if (st) return st;
    *ok = true;
    return HBP_OK;
}

int umma_plan_create(hbp_ctx* ctx, HrnetModel& m, int op_index, int capP, UmmaPlan** out) {
    const HOp& op = m.ops[op_index];
    const HTensor& ti = m.tensors[op.in];
    const int Ho = ti.h / op.stride, Wo = ti.w / op.stride;
    UmmaPlan* pl = new UmmaPlan();
    ConvParams& p = pl->prm;
    memset(&p, 0, sizeof(p));
    EncodeTiledFn enc = get_encode();
    if (!enc) { delete pl; hbp_set_error("cuTensorMapEncodeTiled unavailable"); return HBP_ERR_CUDA; }
    if (!(ctx->attr_flags & ATTR_UMMA)) {
        HBP_CUDA(cudaFuncSetAttribute(conv_umma_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 220 * 1024));
        HBP_CUDA(cudaFuncSetAttribute(conv_umma_halo_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 220 * 1024));
        ctx->attr_flags |= ATTR_UMMA;
    }
    if (op.k == 3 && op.stride == 1 && op.up == 1 && halo_mode_enabled()) {
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
    if (op.k == 1) while (n_tile > 128 && n_tile % 32 == 0) n_tile /= 2;
    // M tiles per CTA: 2 when that still leaves >= 2 CTAs per SM and TMEM stays <= 256 columns
    int m_tiles = 2, tn = 0, th = 0, tw = 0;
    {
        bool ok2 = pick_tile(Ho, Wo, 2, &tn, &th, &tw) && 2 * n_tile <= (op.k == 1 ? 128 : 256);
        if (ok2) {
            const long ctas = (long)((capP + tn - 1) / tn) * (Ho / th) * (Wo / tw) * (op.cout / n_tile);
            if (ctas < 2L * ctx->sm_count) ok2 = false;
        }
        if (!ok2) { m_tiles = 1; pick_tile(Ho, Wo, 1, &tn, &th, &tw); }
    }
    // too few CTAs: split N further (down to 32)
    {
        long ctas = (long)((capP + tn - 1) / tn) * (Ho / th) * (Wo / tw) * (op.cout / n_tile);
        while (ctas < ctx->sm_count && n_tile % 32 == 0 && n_tile > 32) { n_tile /= 2; ctas *= 2; }
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
    {
        int st = encode_weights_map(enc, pl, m, op, p.chunk, n_tile, sw);
        if (st) { delete pl; return st; }
    }
    *out = pl;
    return HBP_OK;
}

int umma_launch(hbp_ctx* ctx, HrnetModel& m, int op_index, UmmaPlan* pl, int P, cudaStream_t st) {
    (void)m; (void)op_index;
    ConvParams p = pl->prm;
    p.P = P;
    const int tiles_n = (P + p.tn - 1) / p.tn;
    dim3 grid((unsigned)(tiles_n * p.tiles_h * p.tiles_w), (unsigned)pl->n_splits);
    if (p.mode == 1) {
        // persistent CTAs: one full wave, every CTA strides over the tiles
        p.n_tiles = (int)grid.x;
        int slots = ctx->sm_count * pl->occ / pl->n_splits;
        if (slots < 1) slots = 1;
        if ((int)grid.x > slots) {
            // equalise: every CTA walks ceil(n_tiles / ctas) or one fewer tiles
            const int per = (p.n_tiles + slots - 1) / slots;
            grid.x = (unsigned)((p.n_tiles + per - 1) / per);
        }
    }
    static const bool trace = getenv("HBP_CONV_TRACE") != nullptr;
    if (trace && p.mode == 1) {
        static int traced = 0;
        if (traced++ % 8 == 3) {        // a warm launch of every shape the process runs
            long long* d = nullptr;
            cudaMalloc(&d, 64 * 32 * sizeof(long long));
            cudaMemset(d, 0, 64 * 32 * sizeof(long long));
            p.dbg = d;
            conv_umma_halo_kernel<<<grid, kThreads, pl->smem_bytes, st>>>(pl->tmA, pl->tmB, p);
            cudaStreamSynchronize(st);
            static long long h[64 * 32];
            cudaMemcpy(h, d, sizeof(h), cudaMemcpyDeviceToHost);
            cudaFree(d);
            const int n = grid.x < 64 ? (int)grid.x : 64;
            double acc[32] = {0};
            for (int i = 0; i < n; ++i) for (int k = 1; k < 32; ++k) acc[k] += (double)(h[i * 32 + k] - h[i * 32]);
            fprintf(stderr, "[taps] B landed / MMAs issued per tap (cycles from CTA start):");
            for (int k = 0; k < 9; ++k) fprintf(stderr, " %.0f/%.0f", acc[8 + k] / n, acc[20 + k] / n);
            fprintf(stderr, "\n");
            fprintf(stderr, "[trace] grid=(%u,%u) smem=%zu stages=%d m=%d n_tile=%d chunks=%d | cycles from CTA start: setup %.0f, A0 landed %.0f, "
                    "B0 landed %.0f, MMAs issued %.0f, accum ready %.0f, epilogue done %.0f, dealloc %.0f\n", grid.x, grid.y, pl->smem_bytes,
                    p.stages, p.m_tiles, p.n_tile, p.n_chunks, acc[1] / n, acc[2] / n, acc[3] / n, acc[4] / n, acc[5] / n, acc[6] / n, acc[7] / n);
            p.dbg = nullptr;
            return HBP_OK;
        }
    }
    if (p.mode == 1) conv_umma_halo_kernel<<<grid, kThreads, pl->smem_bytes, st>>>(pl->tmA, pl->tmB, p);
    else conv_umma_kernel<<<grid, kThreads, pl->smem_bytes, st>>>(pl->tmA, pl->tmB, p);
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return hbp_cuda_fail(e, "conv_umma_kernel", __FILE__, __LINE__);
    return HBP_OK;
}
