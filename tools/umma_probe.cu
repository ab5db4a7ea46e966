// Probe: how does tcgen05.mma address a swizzled K-major shared-memory operand whose
// start address / 8-row-group stride are NOT aligned to the swizzle pattern?
//
// A (R rows x K halfs, K*2 = 128 B or 64 B) is TMA-loaded into smem with the hardware
// swizzle, B = identity (N = K), so D = A[rows read] and the output shows exactly which
// smem rows the MMA fetched for each descriptor variant:
//     MMA row r  ->  smem row  shift + (r/8)*group_rows + (r%8)
// Variants: shift in {0,1,2,3,9,10,11}, group_rows in {8,10,16}, base_offset in {0, auto}.
//
//   nvcc -gencode arch=compute_100a,code=sm_100a -O2 -o umma_probe tools/umma_probe.cu && ./umma_probe
#include <cuda.h>
#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*,
                                  CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion,
                                  CUtensorMapFloatOOBfill);

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

struct Variant { int shift, group_rows, base_mode; };

__global__ void __launch_bounds__(128)
probe_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB, int K, int rows_loaded,
             Variant v, float* out) {
    extern __shared__ __align__(1024) uint8_t smem_raw[];
    const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
    const uint32_t row_bytes = K * 2;
    const uint32_t a_s = base, b_s = base + rows_loaded * row_bytes;     // 1024-aligned (rows_loaded % 16 == 0)
    const uint32_t bar = b_s + 64 * 128, bar2 = bar + 8, slot = bar + 16;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (threadIdx.x == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(bar));
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(bar2));
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(slot), "r"(64) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    uint32_t tmem;
    asm volatile("ld.shared.b32 %0, [%1];" : "=r"(tmem) : "r"(slot));
    if (threadIdx.x == 0) {
        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(rows_loaded * row_bytes + K * row_bytes) : "memory");
        asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
                     ::"r"(a_s), "l"(&tmA), "r"(bar), "r"(0), "r"(0) : "memory");
        asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
                     ::"r"(b_s), "l"(&tmB), "r"(bar), "r"(0), "r"(0) : "memory");
        mbar_wait(bar, 0);
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        const uint32_t idesc = (1u << 4) | ((uint32_t)(K >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
        for (int k = 0; k < K / 16; ++k) {
            const uint32_t a_addr = a_s + v.shift * row_bytes + k * 32;
            uint64_t ad = 0, bd = 0;
            const uint64_t layout = row_bytes == 128 ? 2 : 4;
            ad |= (uint64_t)((a_addr & 0x3FFFF) >> 4);
            ad |= (uint64_t)1 << 16;
            ad |= (uint64_t)((v.group_rows * row_bytes) >> 4) << 32;
            ad |= (uint64_t)1 << 46;
            if (v.base_mode == 1) ad |= (uint64_t)((a_addr >> 7) & 7) << 49;
            ad |= layout << 61;
            bd |= (uint64_t)(((b_s + k * 32) & 0x3FFFF) >> 4);
            bd |= (uint64_t)1 << 16;
            bd |= (uint64_t)((8 * row_bytes) >> 4) << 32;
            bd |= (uint64_t)1 << 46;
            bd |= layout << 61;
            asm volatile("{\n.reg .pred p;\nsetp.ne.b32 p, %4, 0;\ntcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n}\n"
                         ::"r"(tmem), "l"(ad), "l"(bd), "r"(idesc), "r"(k ? 1u : 0u) : "memory");
        }
        asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar2) : "memory");
    }
    mbar_wait(bar2, 0);
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    for (int c0 = 0; c0 < K; c0 += 16) {
        uint32_t r[16];
        asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
                     : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
                       "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
                     : "r"(tmem + ((uint32_t)(warp * 32) << 16) + c0) : "memory");
        asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
        for (int q = 0; q < 16; ++q) out[(warp * 32 + lane) * K + c0 + q] = __uint_as_float(r[q]);
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(64) : "memory");
}

int main() {
    void* fp = nullptr;
    cudaDriverEntryPointQueryResult q;
    cudaFree(0);
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fp, cudaEnableDefault, &q) != cudaSuccess || !fp) { printf("no encode fn\n"); return 1; }
    EncodeTiledFn enc = (EncodeTiledFn)fp;
    const int rows_loaded = 256;           // TMA box rows (max 256)
    int fails = 0;
    for (int K : {64, 32}) {
        std::vector<__half> hA((size_t)rows_loaded * K), hB((size_t)K * K);
        // A[r][c] = r + 256*(c/8): identifies the row and the 16-byte chunk (integers < 2048 are exact in fp16)
        for (int r = 0; r < rows_loaded; ++r) for (int c = 0; c < K; ++c) hA[(size_t)r * K + c] = __float2half((float)(r + 256 * (c / 8)));
        for (int n = 0; n < K; ++n) for (int c = 0; c < K; ++c) hB[(size_t)n * K + c] = __float2half(n == c ? 1.f : 0.f);
        __half *dA, *dB; float* dO;
        cudaMalloc(&dA, hA.size() * 2); cudaMalloc(&dB, hB.size() * 2); cudaMalloc(&dO, 128 * K * 4);
        cudaMemcpy(dA, hA.data(), hA.size() * 2, cudaMemcpyHostToDevice);
        cudaMemcpy(dB, hB.data(), hB.size() * 2, cudaMemcpyHostToDevice);
        CUtensorMap tmA, tmB;
        const CUtensorMapSwizzle sw = K == 64 ? CU_TENSOR_MAP_SWIZZLE_128B : CU_TENSOR_MAP_SWIZZLE_64B;
        cuuint64_t gd[2] = {(cuuint64_t)K, (cuuint64_t)rows_loaded}; cuuint64_t gs[1] = {(cuuint64_t)K * 2};
        cuuint32_t bx[2] = {(cuuint32_t)K, (cuuint32_t)rows_loaded}; cuuint32_t es[2] = {1, 1};
        if (enc(&tmA, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 2, dA, gd, gs, bx, es, CU_TENSOR_MAP_INTERLEAVE_NONE, sw,
                CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE)) { printf("encode A failed\n"); return 1; }
        cuuint64_t gdb[2] = {(cuuint64_t)K, (cuuint64_t)K}; cuuint32_t bxb[2] = {(cuuint32_t)K, (cuuint32_t)K};
        if (enc(&tmB, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 2, dB, gdb, gs, bxb, es, CU_TENSOR_MAP_INTERLEAVE_NONE, sw,
                CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE)) { printf("encode B failed\n"); return 1; }
        const size_t smem = (size_t)rows_loaded * K * 2 + 64 * 128 + 64 + 1024;
        cudaFuncSetAttribute(probe_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        std::vector<float> hO(128 * K);
        for (int group_rows : {8, 10, 16, 18}) for (int shift : {0, 1, 2, 3, 9, 10, 11, 20}) for (int base_mode : {0, 1}) {
            if (shift + 15 * group_rows + 8 > rows_loaded) continue;
            Variant v{shift, group_rows, base_mode};
            cudaMemset(dO, 0, 128 * K * 4);
            probe_kernel<<<1, 128, smem>>>(tmA, tmB, K, rows_loaded, v, dO);
            cudaError_t e = cudaDeviceSynchronize();
            if (e != cudaSuccess) { printf("K=%d variant shift=%d group=%d base=%d: CUDA error %s\n", K, shift, group_rows, base_mode, cudaGetErrorString(e)); return 2; }
            cudaMemcpy(hO.data(), dO, 128 * K * 4, cudaMemcpyDeviceToHost);
            int bad = 0, first_bad = -1;
            for (int r = 0; r < 128; ++r) {
                const int src = shift + (r / 8) * group_rows + (r % 8);
                for (int c = 0; c < K; ++c) {
                    const float want = (float)(src + 256 * (c / 8));
                    if (hO[r * K + c] != want) { if (first_bad < 0) first_bad = r * K + c; ++bad; }
                }
            }
            printf("K=%d (SW%d) group_rows=%2d shift=%2d base_offset=%s : %s", K, K * 2, group_rows, shift,
                   base_mode ? "auto" : "0   ", bad ? "MISMATCH" : "ok");
            if (bad) printf(" (%d wrong, first at r=%d c=%d got %.0f = row %d chunk %d)", bad, first_bad / K, first_bad % K, hO[first_bad], ((int)hO[first_bad]) % 256, ((int)hO[first_bad]) / 256);
            printf("\n");
            if (bad && group_rows == 8 && shift == 0) ++fails;
        }
        cudaFree(dA); cudaFree(dB); cudaFree(dO);
    }
    return fails ? 3 : 0;
}
