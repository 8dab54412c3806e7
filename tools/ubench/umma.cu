// tcgen05.mma dispatch-rate calibration: M=128, K=16, fp16, A from TMEM, B from smem (no swizzle), N in {16,32,64,128,256}.
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__global__ void k_umma(int n_tok, int iters, int ctas_issue_d2, long long* out) {
    extern __shared__ __align__(1024) uint8_t smem[];
    __shared__ uint32_t tmem_slot;
    __shared__ uint64_t bar;
    const int tid = threadIdx.x, warp = tid >> 5;
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_slot)), "r"(512));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
    }
    if (tid == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&bar)));
        asm volatile("fence.mbarrier_init.release.cluster;");
    }
    for (int i = tid; i < 16384 / 4; i += blockDim.x) ((uint32_t*)smem)[i] = 0x3C003C00u;
    asm volatile("fence.proxy.async.shared::cta;");
    asm volatile("tcgen05.fence::before_thread_sync;");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;");
    const uint32_t tb = tmem_slot;
    const uint32_t idesc = (1u << 4) | ((uint32_t)(n_tok >> 3) << 17) | ((128u >> 4) << 24);
    const uint64_t LBO = 128, SBO = 0;
    const uint64_t bdesc = (uint64_t)((smem_u32(smem) >> 4) & 0x3FFF) | ((LBO >> 4) << 16) | ((SBO >> 4) << 32) | (1ull << 46);
    long long t0 = 0, t1 = 0;
    if (warp == 0) {
        uint32_t el = 0;
        asm volatile("{\n .reg .pred px;\n elect.sync _|px, 0xFFFFFFFF;\n @px mov.s32 %0, 1;\n}" : "+r"(el));
        t0 = clock64();
        if (el) {
            for (int it = 0; it < iters; ++it) {
                asm volatile("{\n .reg .pred p;\n setp.ne.b32 p, %4, 0;\n tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, {%5,%5,%5,%5}, p;\n}"
                             ::"r"(tb + 256), "r"(tb + (it & 7) * 8), "l"(bdesc + (uint64_t)((it & 7) * 16)), "r"(idesc), "r"(1u), "r"(0u));
                if (ctas_issue_d2)
                    asm volatile("{\n .reg .pred p;\n setp.ne.b32 p, %4, 0;\n tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, {%5,%5,%5,%5}, p;\n}"
                                 ::"r"(tb + 256 + 16), "r"(tb + 64), "l"(bdesc + (uint64_t)((it & 7) * 16)), "r"(idesc), "r"(1u), "r"(0u));
            }
            asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&bar)));
        }
        __syncwarp();
        t1 = clock64();
        asm volatile("{\n .reg .pred p;\n W: mbarrier.try_wait.parity.shared::cta.b64 p, [%0], 0;\n @p bra D;\n bra W;\n D:\n}" ::"r"(smem_u32(&bar)));
        long long t2 = clock64();
        if (tid == 0 && blockIdx.x == 0) { out[0] = t1 - t0; out[1] = t2 - t0; }
    }
    asm volatile("tcgen05.fence::before_thread_sync;");
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tb), "r"(512));
}

int main() {
    long long* out; cudaMallocManaged(&out, 64);
    cudaFuncSetAttribute(k_umma, cudaFuncAttributeMaxDynamicSharedMemorySize, 32768);
    const int iters = 4096;
    for (int d2 = 0; d2 < 2; ++d2)
    for (int n : {16, 32, 64, 128, 256}) {
        k_umma<<<148, 128, 32768>>>(n, iters, d2, out); cudaDeviceSynchronize();
        k_umma<<<148, 128, 32768>>>(n, iters, d2, out);
        cudaError_t e = cudaDeviceSynchronize();
        int nm = iters * (d2 ? 2 : 1);
        printf("N=%3d d2=%d: issue %.1f clk/MMA, complete %.1f clk/MMA  (%s)\n", n, d2, (double)out[0] / nm, (double)out[1] / nm, cudaGetErrorString(e));
    }
    return 0;
}
