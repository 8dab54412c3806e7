// Latency of tcgen05.st (+wait::st) and tcgen05.ld (+wait::ld) as seen by one warp, alone and with other warps busy.
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tmem_lat tmem_lat.cu && ./tmem_lat
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__global__ void k(int iters, int nst, long long* out) {
    __shared__ uint32_t slot;
    const int warp = threadIdx.x >> 5;
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&slot)), "r"(256));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
    }
    asm volatile("tcgen05.fence::before_thread_sync;");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;");
    const uint32_t t = slot + (((uint32_t)(warp & 3) * 32) << 16) + (warp >> 2) * 64;
    uint32_t r[16];
    for (int i = 0; i < 16; ++i) r[i] = threadIdx.x * 17 + i;
    long long st_sum = 0, ld_sum = 0;
    for (int it = 0; it < iters; ++it) {
        long long t0 = clock64();
        for (int s = 0; s < nst; ++s)
            asm volatile("tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16};" ::"r"(t + s * 16),
                         "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]), "r"(r[8]), "r"(r[9]), "r"(r[10]), "r"(r[11]),
                         "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15]));
        asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
        long long t1 = clock64();
        uint32_t v;
        asm volatile("tcgen05.ld.sync.aligned.32x32b.x1.b32 {%0}, [%1];" : "=r"(v) : "r"(t) : "memory");
        asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
        long long t2 = clock64();
        r[0] ^= v;
        st_sum += t1 - t0;
        ld_sum += t2 - t1;
    }
    if ((threadIdx.x & 31) == 0 && blockIdx.x == 0) {
        out[warp * 2] = st_sum / iters;
        out[warp * 2 + 1] = ld_sum / iters;
    }
    asm volatile("tcgen05.fence::before_thread_sync;");
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(slot), "r"(256));
    if (r[0] == 0x1234567) out[63] = r[0];
}
int main() {
    long long* out;
    cudaMallocManaged(&out, 64 * 8);
    for (int warps : {1, 4, 8}) for (int nst : {1, 2, 4}) {
        for (int ctas : {1, 296}) {
            k<<<ctas, warps * 32>>>(2000, nst, out);
            cudaDeviceSynchronize();
            printf("warps %d ctas %3d  %d x STTM.x16 + wait::st = %lld clk   LDTM.x1 + wait::ld = %lld clk  (%s)\n", warps, ctas, nst, out[0], out[1],
                   cudaGetErrorString(cudaGetLastError()));
        }
    }
    return 0;
}
