// Probe: cost of waiting on an mbarrier phase that has already completed (cycles per call, one warp).
//   nvcc -gencode arch=compute_100a,code=sm_100a -O2 -o tools/mbar_probe tools/mbar_probe.cu
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdint>
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
template <int MODE>
__device__ __forceinline__ uint32_t wait_once(uint32_t bar, uint32_t parity) {
    uint32_t done;
    if (MODE == 0)
        asm volatile("{\n.reg .pred p;\nmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\nselp.u32 %0, 1, 0, p;\n}\n" : "=r"(done) : "r"(bar), "r"(parity) : "memory");
    else if (MODE == 1)
        asm volatile("{\n.reg .pred p;\nmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2, %3;\nselp.u32 %0, 1, 0, p;\n}\n" : "=r"(done) : "r"(bar), "r"(parity), "r"(0x989680u) : "memory");
    else
        asm volatile("{\n.reg .pred p;\nmbarrier.test_wait.parity.shared::cta.b64 p, [%1], %2;\nselp.u32 %0, 1, 0, p;\n}\n" : "=r"(done) : "r"(bar), "r"(parity) : "memory");
    return done;
}
template <int MODE>
__global__ void k(long long* out, int fence) {
    __shared__ uint64_t bars[4];
    const uint32_t bar = smem_u32(&bars[0]);
    if (threadIdx.x == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(bar));
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");     // phase 0 complete
    }
    __syncthreads();
    if (threadIdx.x < 32) {
        uint32_t acc = 0;
        const long long t0 = clock64();
        for (int i = 0; i < 64; ++i) {
            while (!wait_once<MODE>(bar, 0)) {}
            if (fence) asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
            acc += i;
        }
        const long long t1 = clock64();
        if (threadIdx.x == 0) { out[0] = t1 - t0; out[1] = acc; }
    }
}
int main() {
    long long* d; cudaMalloc(&d, 16);
    long long h[2];
    for (int fence = 0; fence < 2; ++fence) {
        k<0><<<1, 128>>>(d, fence); cudaMemcpy(h, d, 16, cudaMemcpyDeviceToHost); printf("try_wait            fence=%d: %.1f cycles/wait\n", fence, h[0] / 64.0);
        k<1><<<1, 128>>>(d, fence); cudaMemcpy(h, d, 16, cudaMemcpyDeviceToHost); printf("try_wait + timehint fence=%d: %.1f cycles/wait\n", fence, h[0] / 64.0);
        k<2><<<1, 128>>>(d, fence); cudaMemcpy(h, d, 16, cudaMemcpyDeviceToHost); printf("test_wait           fence=%d: %.1f cycles/wait\n", fence, h[0] / 64.0);
    }
    return 0;
}
