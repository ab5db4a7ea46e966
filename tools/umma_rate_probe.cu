// Probe: issue rate and execution rate of small tcgen05.mma (M = 128, K = 16, fp16) as a
// function of N, of the number of issuing warps per CTA and of the number of independent
// accumulators each warp rotates over.  Operand contents are irrelevant (uninitialised
// shared memory); only the clocks matter.
//
//   nvcc -gencode arch=compute_100a,code=sm_100a -O2 -o tools/umma_rate_probe tools/umma_rate_probe.cu
//   ./tools/umma_rate_probe
//
// Per configuration it prints, per MMA and averaged over the issuing warps of CTA 0:
//   issue  = cycles between the first tcgen05.mma and the instruction after the last one
//   exec   = cycles until the tcgen05.commit barrier of the warp's MMAs completes
//   sm     = cycles per MMA seen by the SM (max over warps of exec time / total MMAs of the CTA)
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdlib>
#include <cstdint>
#include <vector>

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
    const long long t0 = clock64();
    for (;;) {
        uint32_t done;
        asm volatile("{\n.reg .pred p;\nmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\nselp.u32 %0, 1, 0, p;\n}\n"
                     : "=r"(done) : "r"(bar), "r"(parity) : "memory");
        if (done) return;
        if (clock64() - t0 > 2000000000LL) __trap();
    }
}
__device__ __forceinline__ bool elect_one() {
    uint32_t pred;
    asm volatile("{\n.reg .pred p;\nelect.sync _|p, 0xffffffff;\nselp.u32 %0, 1, 0, p;\n}\n" : "=r"(pred));
    return pred != 0;
}
__device__ __forceinline__ void umma(uint32_t d, uint64_t a, uint64_t b, uint32_t idesc, uint32_t acc) {
    asm volatile("{\n.reg .pred p;\nsetp.ne.b32 p, %4, 0;\ntcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n}\n"
                 ::"r"(d), "l"(a), "l"(b), "r"(idesc), "r"(acc) : "memory");
}
__device__ __forceinline__ uint64_t make_desc(uint32_t saddr, uint32_t row_bytes, uint32_t sbo_bytes) {
    uint64_t d = 0;
    d |= (uint64_t)((saddr & 0x3FFFF) >> 4);
    d |= (uint64_t)1 << 16;
    d |= (uint64_t)(sbo_bytes >> 4) << 32;
    d |= (uint64_t)1 << 46;
    d |= (uint64_t)(row_bytes == 128 ? 2 : 4) << 61;
    return d;
}

struct Cfg { int N, row_bytes, warps, chains, n_mma, group_rows, distinct_a, mode; float* scratch; };
// mode bits: 1 = idle warps spin on an mbarrier that never completes (try_wait loop), 2 = idle warps run tcgen05.ld loops,
//            4 = the issue loop is cut into slots of 6 MMAs with fence / elect / syncwarp around each, 8 = idle warps stream global memory

template <int CHAINS>
__global__ void __launch_bounds__(256)
rate_kernel(Cfg c, long long* out) {
    extern __shared__ __align__(1024) uint8_t smem_raw[];
    const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
    const uint32_t a_s = base;                       // 64 KB of "A"
    const uint32_t b_s = base + 64 * 1024;           // 32 KB of "B"
    const uint32_t bar = b_s + 32 * 1024, slot = bar + 64;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (threadIdx.x == 0) {
        for (int i = 0; i < 8; ++i) asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(bar + 8 * i));
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(slot), "r"(512) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    uint32_t tmem;
    asm volatile("ld.shared.b32 %0, [%1];" : "=r"(tmem) : "r"(slot));
    tmem = __shfl_sync(0xffffffffu, tmem, 0);
    const int w = __shfl_sync(0xffffffffu, warp, 0);
    if (w < c.warps) {
        const uint32_t idesc = (1u << 4) | ((uint32_t)(c.N >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
        const uint64_t ad0 = make_desc(a_s + (uint32_t)w * 8192u, c.row_bytes, c.group_rows * c.row_bytes);
        const uint64_t bd0 = make_desc(b_s, c.row_bytes, 8 * c.row_bytes);
        const uint32_t d0 = tmem + (uint32_t)(w * CHAINS * c.N);
        const uint32_t astep = c.distinct_a ? (uint32_t)(c.row_bytes >> 4) : 0u;    // shift A by one pixel row per chain
        long long t0 = 0, t1 = 0, t2 = 0;
        __syncwarp();
        t0 = clock64();
        if (c.mode & 4) {
            for (int i = 0; i < c.n_mma; i += 6) {
                asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                if (elect_one()) {
                    for (int k = 0; k < 6; ++k)
                        umma(d0 + (k % CHAINS) * c.N, ad0 + k * astep, bd0 + 2 * (k & 1), idesc, (i | k) ? 1u : 0u);
                }
                __syncwarp();
            }
            if (elect_one())
                asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar + 8 * w) : "memory");
            __syncwarp();
        } else {
        if (elect_one()) {
            for (int i = 0; i < c.n_mma; i += CHAINS) {
#pragma unroll
                for (int k = 0; k < CHAINS; ++k)
                    umma(d0 + k * c.N, ad0 + k * astep, bd0 + 2 * (k & 1), idesc, i ? 1u : 0u);
            }
            asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar + 8 * w) : "memory");
        }
        __syncwarp();
        }
        t1 = clock64();
        mbar_wait(bar + 8 * w, 0);
        t2 = clock64();
        if (lane == 0 && blockIdx.x == 0) { out[w * 2] = t1 - t0; out[w * 2 + 1] = t2 - t0; }
    }
    else {
        // "idle" warps: emulate what the other roles of a real kernel do while the MMA warp issues.
        // They stop when the issuing warp 0 has finished (its commit barrier, parity 0, completes).
        if (c.mode & 16) {
            // fixed amount of dependent-free ALU work: how long does it take on each scheduler while warp 0 issues MMAs?
            float a0 = threadIdx.x, a1 = 1.f, a2 = 2.f, a3 = 3.f;
            const long long s0 = clock64();
            for (int it = 0; it < 2000; ++it) {
                a0 = fmaf(a0, 1.0001f, 0.5f); a1 = fmaf(a1, 1.0001f, 0.5f); a2 = fmaf(a2, 1.0001f, 0.5f); a3 = fmaf(a3, 1.0001f, 0.5f);
            }
            const long long s1 = clock64();
            if (a0 + a1 + a2 + a3 == 123.f) c.scratch[0] = a0;
            if (lane == 0 && blockIdx.x == 0) out[16 + warp] = s1 - s0;
        } else if (c.mode & 1) {
            mbar_wait(bar + 8 * 0, 0);
        } else if (c.mode & 2) {
            float acc = 0.f;
            for (int it = 0; it < 4000; ++it) {
                uint32_t r[16];
                asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
                             : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
                               "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
                             : "r"(tmem + ((uint32_t)((warp & 3) * 32) << 16) + 256) : "memory");
                asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
                acc += __uint_as_float(r[it & 15]);
                uint32_t done;
                asm volatile("{\n.reg .pred p;\nmbarrier.test_wait.parity.shared::cta.b64 p, [%1], %2;\nselp.u32 %0, 1, 0, p;\n}\n"
                             : "=r"(done) : "r"(bar), "r"(0) : "memory");
                if (done) break;
            }
            if (acc == 123.f) c.scratch[0] = acc;
        } else if (c.mode & 8) {
            float acc = 0.f;
            for (int it = 0; it < 4000; ++it) {
                const float4 v = *reinterpret_cast<const float4*>(c.scratch + ((size_t)blockIdx.x * 4096 + (it & 31) * 1024 + threadIdx.x * 4));
                acc += v.x + v.y;
                *reinterpret_cast<float4*>(c.scratch + (1 << 22) + ((size_t)blockIdx.x * 4096 + (it & 31) * 1024 + threadIdx.x * 4)) = v;
                uint32_t done;
                asm volatile("{\n.reg .pred p;\nmbarrier.test_wait.parity.shared::cta.b64 p, [%1], %2;\nselp.u32 %0, 1, 0, p;\n}\n"
                             : "=r"(done) : "r"(bar), "r"(0) : "memory");
                if (done) break;
            }
            if (acc == 123.f) c.scratch[0] = acc;
        }
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(512) : "memory");
}

template <int CHAINS>
static void run(const Cfg& c, int grid, long long* d_out) {
    const size_t smem = 64 * 1024 + 32 * 1024 + 256 + 1024;
    cudaFuncSetAttribute(rate_kernel<CHAINS>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    rate_kernel<CHAINS><<<grid, 256, smem>>>(c, d_out);      // warm
    rate_kernel<CHAINS><<<grid, 256, smem>>>(c, d_out);
}

int main() {
    long long* d_out;
    cudaMalloc(&d_out, 64 * sizeof(long long));
    int dev_sms = 0;
    cudaDeviceGetAttribute(&dev_sms, cudaDevAttrMultiProcessorCount, 0);
    printf("%4s %4s %5s %6s %6s %5s %6s | %8s %8s %8s\n", "N", "rowB", "warps", "chains", "mode", "dA", "grid", "issue/MMA", "exec/MMA", "sm/MMA");
    float* scratch;
    cudaMalloc(&scratch, (size_t)(1 << 23) * sizeof(float) + (size_t)dev_sms * 4096 * 32 * 4);
    cudaMemset(scratch, 0, (size_t)(1 << 23) * sizeof(float));
    for (int grid : {1, dev_sms})
        for (int N : {32, 64, 128, 256})
            for (int row_bytes : {64, 128})
                for (int warps : {1, 2, 4})
                    for (int chains : {1, 4})
                        for (int mode : {0, 16}) {
                            const int group_rows = 10, dA = 1;
                            if (warps * chains * N > 512) continue;
                            if (row_bytes == 64 && N > 64) continue;
                            if (row_bytes == 128 && N == 32) continue;
                            if (mode && (warps > 1 || chains > 1)) continue;
                            if (grid > 1 && warps == 2) continue;
                            Cfg c{N, row_bytes, warps, chains, mode ? 1008 : 252, group_rows, dA, mode, scratch};
                            cudaMemset(d_out, 0, 64 * sizeof(long long));
                            if (chains == 1) run<1>(c, grid, d_out);
                            else run<4>(c, grid, d_out);
                            cudaError_t e = cudaDeviceSynchronize();
                            if (e != cudaSuccess) { printf("CUDA error %s\n", cudaGetErrorString(e)); return 2; }
                            long long h[64];
                            cudaMemcpy(h, d_out, sizeof(h), cudaMemcpyDeviceToHost);
                            double issue = 0, exec = 0, mx = 0;
                            for (int w = 0; w < warps; ++w) {
                                issue += (double)h[2 * w] / c.n_mma;
                                exec += (double)h[2 * w + 1] / c.n_mma;
                                if ((double)h[2 * w + 1] > mx) mx = (double)h[2 * w + 1];
                            }
                            printf("%4d %4d %5d %6d %6d %5d %6d | %8.1f %8.1f %8.1f", N, row_bytes, warps, chains, mode, dA, grid,
                                   issue / warps, exec / warps, mx / (c.n_mma * warps));
                            if (mode & 16) { printf("  | ALU loop cycles, warps 1..7:"); for (int w = 1; w < 8; ++w) printf(" %lld", h[16 + w]); }
                            printf("\n");
                        }
    cudaFree(d_out);
    return 0;
}
