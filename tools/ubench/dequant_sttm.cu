// Throughput of the 4-bit unpack -> tcgen05.st path (no MMA): slabs (32 k x 32 n = 1024 weights per warp-iteration).
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
#include "../../exllamav2_b200/csrc/dequant.cuh"
using namespace exl2b;
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void st16(uint32_t taddr, const uint32_t* r, bool vol) {
    asm volatile("tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16};"
        ::"r"(taddr), "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]), "r"(r[8]),
        "r"(r[9]), "r"(r[10]), "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15]));
}
template <int MODE>   // 0: unpack only (xor-reduce), 1: unpack + st, 2: unpack + st + wait every 4
__global__ void k(int iters, long long* out, uint32_t* sink) {
    extern __shared__ __align__(1024) uint8_t smem[];
    __shared__ uint32_t tmem_slot;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_slot)), "r"(512));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
    }
    for (int i = tid; i < 32768 / 4; i += blockDim.x) ((uint32_t*)smem)[i] = i * 2654435761u;
    asm volatile("tcgen05.fence::before_thread_sync;");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;");
    const uint32_t tb = tmem_slot + ((uint32_t)((warp & 3) * 32) << 16) + (warp >> 2) * 64;
    uint32_t acc = 0;
    long long t0 = clock64();
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            const uint4 a = *reinterpret_cast<const uint4*>(smem + ((it * 4 + i) & 31) * 512 + (warp & 1) * 16384 + lane * 16);
            uint32_t mw[4] = {a.x, a.y, a.z, a.w}, A[16];
            dequant_block_4bit_offset(mw, A);
            if (MODE == 0) { for (int j = 0; j < 16; ++j) acc ^= A[j]; }
            else st16(tb + i * 16, A, true);
        }
        if (MODE == 2) asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
    }
    if (MODE == 1) asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
    long long t1 = clock64();
    if (tid == 0 && blockIdx.x == 0) out[0] = t1 - t0;
    if (acc == 0x12345) sink[tid] = acc;
    asm volatile("tcgen05.fence::before_thread_sync;");
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_slot), "r"(512));
}
int main() {
    long long* out; cudaMallocManaged(&out, 64);
    uint32_t* sink; cudaMalloc(&sink, 4096);
    const int iters = 2048;
    cudaFuncSetAttribute(k<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, 32768);
    cudaFuncSetAttribute(k<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, 32768);
    cudaFuncSetAttribute(k<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, 32768);
    for (int warps : {4, 8, 16}) {
        for (int mode = 0; mode < 3; ++mode) {
            for (int rep = 0; rep < 2; ++rep) {
                if (mode == 0) k<0><<<148, warps * 32, 32768>>>(iters, out, sink);
                if (mode == 1) k<1><<<148, warps * 32, 32768>>>(iters, out, sink);
                if (mode == 2) k<2><<<148, warps * 32, 32768>>>(iters, out, sink);
                cudaDeviceSynchronize();
            }
            double clk_per_group = (double)out[0] / iters;       // per warp: 4 slabs
            printf("warps/SM %2d mode %d: %.0f clk per 4 slabs per warp -> %.1f weights/clk/SM (%s)\n", warps, mode, clk_per_group,
                   4096.0 * warps / clk_per_group, cudaGetErrorString(cudaGetLastError()));
        }
    }
    return 0;
}
