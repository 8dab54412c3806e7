// tcgen05 dequant-GEMM: packed EXL2 / GPTQ weights -> fp16 in TENSOR MEMORY -> tcgen05.mma against the activations.
// Serves batch-1 decode through batched rows with ONE kernel (LAYOUT_TC matrices); replaces gemm_half_q_half_kernel
// (exllamav2_ext/cuda/q_gemm_kernel.cuh:140-565), gemm_half_q_half_gptq_kernel (q_gemm_kernel_gptq.cuh:61-246) and the
// reconstruct + cublasHgemm detour for M > 32 (cuda/q_gemm.cu:233-266).
//
// Why tensor memory even for M = 1: measured on B200 (tools/ubench/pipes.cu, profiles/), the legacy mma.sync path
// tops out near 0.45 HMMA.16816/clk/SM and a dequant+mma.sync loop at ~52 4-bit weights/clk/SM -- the whole SM is
// busy just keeping up with HBM (49 weights/clk/SM at 6.5 TB/s).  With UMMA the weights are the M = 128 operand:
//     D[128 weight columns x 16 tokens] += A[128 x 16 k] (TMEM) * B[16 k x 16 tokens] (shared memory)
// one instruction covers 2048 weights in 8 clk, and the only per-weight work left on the CUDA cores is the unpack
// (3 SHF + 4 LOP3 per 8 four-bit weights) plus one tcgen05.st per 1024 weights.
//
// Mapping (layout.h, LAYOUT_TC): strip = 128 output columns = the 128 TMEM lanes; a thread IS a weight column n.
//   * 256 threads = 2 warpgroups (WG).  Warp w of a WG owns TMEM lane quadrant w and streams ITS 32-column block of the
//     strip with cp.async.bulk (TMA 1-D) into a private 3-stage ring, exactly one quantisation group (<= 128 k) per
//     stage.  WG0 takes even groups, WG1 odd groups: two independent software pipelines sharing the tensor core.
//   * per group: unpack 32 k per LDS.128 into 16 half2 registers = 16 TMEM columns of the thread's row
//     (tcgen05.st.32x32b.x16), barrier inside the WG, one thread issues 2 MMAs per 32 k into the WG's accumulator
//     D[wg] (and, for the 4-bit offset form, 2 more against an all-ones A tile to obtain sum_k a[k] per token),
//     tcgen05.commit -> mbarrier.  One group later the WG reads D back (tcgen05.ld), applies the group's fp16 scale
//     in fp32 -- a per-THREAD scalar, since a thread is a column -- and adds into its running totals.
//   * activations: the 128 threads of a WG write the group's [128 k x 16 token] B tile (no-swizzle K-major core
//     matrices) into shared memory, gathered through q_perm, RMSNorm folded in.
//   * epilogue: WG0 + WG1 totals are combined through shared memory; split-K across CTAs and the bias / residual /
//     silu(gate)*up epilogues are the ones of the mma.sync kernel (fixed-order, deterministic).
#include <algorithm>
#include <mutex>
#include <type_traits>

#include "dequant.cuh"
#include "gemv.cuh"

namespace exl2b {

constexpr int TC_THREADS = 320;                   // 8 unpack warps (2 warpgroups) + 2 MMA-issue warps (one per warpgroup)
constexpr int TC_WARPS = 8;
constexpr int TC_MAX_STAGES = 4;                  // stage = one group (<= 4 slabs) of one 32-column block; count + size set per launch
constexpr int TC_NTOK = 16;                       // UMMA N; tokens 8..15 alias tokens 0..7 (SBO = 0), only 8 are real
constexpr int TC_ACT_STAGE = 2048;                // activations of one group: 128 k x 16 B (8 token slots)
constexpr int TC_NBARS = TC_WARPS * TC_MAX_STAGES + 2 * TC_MAX_STAGES + 2 * 3 + 2 * 2 + 2 * 3;   // weights | act | A free | D ready | A full
constexpr int TC_SMEM_BARS = TC_NBARS * 8;
constexpr int TC_SMEM_MISC = 128;                 // tmem base, rstd[8], flag
constexpr int TC_TMEM_COLS = 256;
constexpr int TC_A_BUFS = 3;                      // per WG: 3 A buffers of 32 columns (2 slabs = 64 k each) ...
constexpr int TC_COLS_PER_WG = 128;               //         ... + 2 accumulators of 16 columns
constexpr int TC_RED_FLOATS = GEMV_MTOK * 128;    // workspace floats per (strip, contributor)

// ---- PTX wrappers ----------------------------------------------------------------------------------------------------
__device__ __forceinline__ void tmem_alloc(uint32_t smem_dst, uint32_t cols) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_dst), "r"(cols) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr, uint32_t cols) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(cols) : "memory");
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tmem_wait_st() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }
__device__ __forceinline__ void tmem_wait_ld() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }
__device__ __forceinline__ void fence_proxy_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void mbar_arrive(uint32_t bar) { asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory"); }
__device__ __forceinline__ void bar_sync(int id, int n) { asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(n) : "memory"); }

__device__ __forceinline__ void tmem_st16(uint32_t taddr, const uint32_t* r) {
    asm volatile(
        "tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16};"
        ::"r"(taddr), "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]), "r"(r[8]),
        "r"(r[9]), "r"(r[10]), "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15]));
    // no "memory" clobber: the store touches tensor memory only, so the compiler may overlap the next slab's
    // shared-memory loads and unpack with it (ordering against the MMA is tcgen05.wait::st + fence + barrier)
}
__device__ __forceinline__ void tmem_st8(uint32_t taddr, const uint32_t* r) {
    asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};" ::"r"(taddr), "r"(r[0]), "r"(r[1]),
                 "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7])
                 : "memory");
}
__device__ __forceinline__ void tmem_ld1(uint32_t taddr, float& v) {
    uint32_t r;
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x1.b32 {%0}, [%1];" : "=r"(r) : "r"(taddr) : "memory");
    v = __uint_as_float(r);
}
__device__ __forceinline__ void tmem_ld8(uint32_t taddr, float* v) {
    uint32_t r[8];
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7])
                 : "r"(taddr)
                 : "memory");
#pragma unroll
    for (int i = 0; i < 8; ++i) v[i] = __uint_as_float(r[i]);
}
// D[tmem] (+)= A[tmem] * B[smem desc];  M = 128, N = 16, K = 16, fp16 in, fp32 accumulate
__device__ __forceinline__ void umma_ts(uint32_t d_tmem, uint32_t a_tmem, uint64_t b_desc, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, {%5, %6, %7, %8}, p;\n\t"
        "}\n" ::"r"(d_tmem),
        "r"(a_tmem), "l"(b_desc), "r"(idesc), "r"(accumulate), "r"(0u), "r"(0u), "r"(0u), "r"(0u)
        : "memory");
}
// one lane of a converged warp (the tcgen05.mma / commit issuer); returns non-zero in the elected lane
__device__ __forceinline__ uint32_t elect_one() {
    uint32_t pred = 0;
    asm volatile(
        "{\n\t"
        ".reg .pred px;\n\t"
        "elect.sync _|px, 0xFFFFFFFF;\n\t"
        "@px mov.s32 %0, 1;\n\t"
        "}" : "+r"(pred));
    return pred;
}
__device__ __forceinline__ void umma_commit(uint32_t bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}

// instruction descriptor (cute/arch/mma_sm100_desc.hpp InstrDescriptor): F32 accumulate, F16 x F16, K-major A and B,
// N = 16, M = 128
constexpr uint32_t TC_IDESC = (1u << 4) | ((uint32_t)(TC_NTOK >> 3) << 17) | ((128u >> 4) << 24);
// shared-memory descriptor of a no-swizzle K-major operand: core matrix = 8 rows x 16 bytes, contiguous (128 B);
// LBO = byte distance between core matrices adjacent in K, SBO = between core matrices adjacent in N (tokens)
// The staged activations hold ONE 8-token core matrix per 8 k (128 B); the descriptor's token-group stride is 0, so
// UMMA columns 8..15 re-read tokens 0..7 (their D columns are never looked at).
__device__ __forceinline__ uint64_t make_b_desc(uint32_t smem_byte_addr) {
    constexpr uint64_t LBO = 128, SBO = 0;
    return (uint64_t)((smem_byte_addr >> 4) & 0x3FFF) | ((LBO >> 4) << 16) | ((SBO >> 4) << 32) | (1ull << 46);
}
// byte offset of element (k, tok < 8) in the staged activations (k relative to the segment start)
__device__ __forceinline__ int b_off(int k, int tok) { return (k >> 3) * 128 + tok * 16 + (k & 7) * 2; }

__device__ __forceinline__ int tc_cta_of_unit(unsigned x, unsigned G, unsigned U) { return (int)(((x + 1u) * G - 1u) / U); }
__device__ __forceinline__ int tc_region_of(const QMatView& w, int ks) {
    int r = 0;
#pragma unroll
    for (int i = 1; i < MAX_REGIONS; ++i)
        if (i < w.num_regions && ks >= w.reg[i].ks_begin) r = i;
    return r;
}
__device__ __forceinline__ int tc_region_end(const QMatView& w, int r) { return (r + 1 < w.num_regions) ? w.reg[r + 1].ks_begin : w.KS; }
// first slab of the group that contains slab ks (ks == KS maps to KS)
__device__ __forceinline__ int tc_group_start(const QMatView& w, int ks) {
    if (ks >= w.KS) return w.KS;
    const QRegion& R = w.reg[tc_region_of(w, ks)];
    return R.ks_begin + (((ks - R.ks_begin) >> R.spg_log2) << R.spg_log2);
}

__device__ __forceinline__ half tc_silu_h(half x) {
    half e = hexp(__hneg(x));
    half r = hrcp(__hadd(__float2half(1.0f), e));
    return __hmul(x, r);
}
__device__ __forceinline__ half tc_gelu_h(half x) {
    float xf = __half2float(x);
    const float c = 0.797884560803f;
    float t = c * (xf + 0.044715f * xf * xf * xf), th;
    asm("tanh.approx.f32 %0, %1;" : "=f"(th) : "f"(t));
    xf = 0.5f * xf * (1.0 + th);
    return __float2half_rn(xf);
}

__device__ __forceinline__ unsigned long long tc_gtimer() {
    unsigned long long t;
    asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t));
    return t;
}
#define TC_STAMP(i) do { if (P.dbg) { if (blockIdx.x == P.dbg_cta && tid == 0) P.dbg[i] = tc_gtimer(); if ((i) == 0 && tid == 0) atomicMin(P.dbg + 6, tc_gtimer()); } } while (0)

// -DEXL2B_TC_PROFILE: accumulate clock64 deltas of the main loop's phases in a few warps of CTA dbg_cta (tools/microbench.py --phases)
#ifdef EXL2B_TC_PROFILE
#define TCP_DECL long long tcp_t = 0, tcp_acc[4] = {0, 0, 0, 0}; const bool tcp_on = P.dbg && blockIdx.x == P.dbg_cta;
#define TCP_BEGIN do { if (tcp_on) tcp_t = clock64(); } while (0)
#define TCP_END(i) do { if (tcp_on) { const long long tcp_n = clock64(); tcp_acc[i] += tcp_n - tcp_t; tcp_t = tcp_n; } } while (0)
#define TCP_FLUSH do { if (tcp_on && lane == 0) { const int tcp_slot = warp == 0 ? 0 : warp == 1 ? 1 : warp == 4 ? 2 : warp == 8 ? 3 : -1; \
    if (tcp_slot >= 0) for (int i = 0; i < 4; ++i) P.dbg[8 + tcp_slot * 4 + i] = (unsigned long long)tcp_acc[i]; } } while (0)
#else
#define TCP_DECL
#define TCP_BEGIN do {} while (0)
#define TCP_END(i) do {} while (0)
#define TCP_FLUSH do {} while (0)
#endif

template <int BITS>
__device__ __forceinline__ void tc_load_words(const uint8_t* base, int lane, uint32_t* mw, uint32_t* ew) {
    constexpr int Pm = plane_main(BITS), Pe = plane_extra(BITS);
    if constexpr (Pm == 8) {
        const uint4 a = *reinterpret_cast<const uint4*>(base + lane * 16), b = *reinterpret_cast<const uint4*>(base + 512 + lane * 16);
        mw[0] = a.x; mw[1] = a.y; mw[2] = a.z; mw[3] = a.w; mw[4] = b.x; mw[5] = b.y; mw[6] = b.z; mw[7] = b.w;
    } else if constexpr (Pm == 4) {
        const uint4 a = *reinterpret_cast<const uint4*>(base + lane * 16);
        mw[0] = a.x; mw[1] = a.y; mw[2] = a.z; mw[3] = a.w;
    } else {
        const uint2 a = *reinterpret_cast<const uint2*>(base + lane * 8);
        mw[0] = a.x; mw[1] = a.y;
    }
    if constexpr (Pe == 1) {
        ew[0] = *reinterpret_cast<const uint32_t*>(base + 128 * Pm + lane * 4);
    } else if constexpr (Pe == 2) {
        const uint2 a = *reinterpret_cast<const uint2*>(base + 128 * Pm + lane * 8);
        ew[0] = a.x; ew[1] = a.y;
    }
}

// 4-bit fields in "two-offset form": value = offset + q with offset 1024 (even pair slots) or 64 (odd pair slots), i.e. the
// bare (w & mask) | magic -- 4 LOP3 + 1 SHF per 8 weights and nothing else.  The offsets and the zero point are removed per
// group AFTER the tensor core:  sum_k a_k (q_k - z) = D - (S1 + z * S0),  S1 = sum_k a_k * offset_k,  S0 = sum_k a_k;  S1 and
// S0 do not depend on the output column, the MMA-issue warp computes them once per group from the staged activations.
// Products (offset + q) * a are exact in the tensor core and fp32 accumulation sees operands < 2^11 |a|.
__device__ __forceinline__ void tc_dequant4(const uint32_t* mw, uint32_t* A) {
    const uint32_t m0 = 0x000f000fu, g0 = 0x64006400u, m1 = 0x00f000f0u, g1 = 0x54005400u;
#pragma unroll
    for (int w = 0; w < 4; ++w) {
        const uint32_t x = mw[w], y = x >> 8;
        A[w * 4 + 0] = and_or(x, m0, g0);
        A[w * 4 + 1] = and_or(x, m1, g1);
        A[w * 4 + 2] = and_or(y, m0, g0);
        A[w * 4 + 3] = and_or(y, m1, g1);
    }
}
// unpack slab i of this warp's block stage (contiguous at sp) into 16 TMEM columns at a_taddr
template <int BITS>
__device__ __forceinline__ void tc_dequant_slab(const uint8_t* sp, int i, int lane, uint32_t a_taddr) {
    uint32_t mw[8], ew[2], A[16];
    tc_load_words<BITS>(sp + i * block_bytes(BITS), lane, mw, ew);
    if constexpr (BITS == 4) tc_dequant4(mw, A);
    else dequant_block_exl2<BITS>(mw, ew, A);
    tmem_st16(a_taddr, A);
}
// one chunk = slabs [s0, s0 + nsl) of the stage, nsl <= 2, into one A buffer
template <int BITS>
__device__ __forceinline__ void tc_dequant_chunk(const uint8_t* sp, int s0, int nsl, int lane, uint32_t a_taddr) {
    tc_dequant_slab<BITS>(sp, s0, lane, a_taddr);
    if (nsl == 2) tc_dequant_slab<BITS>(sp, s0 + 1, lane, a_taddr + 16);
}

// ---- activation prep: RMSNorm + q_perm gather + UMMA core-matrix layout, once per launch ------------------------------
// xp[mat][k'/8][tok 0..7][k'%8] = half(x[tok][perm_mat[k']] * w[perm] * rstd[tok])   (rows tok >= M are zero)
// One CTA per token slot (8 CTAs).  The GEMV CTAs then fetch their k-range with ONE bulk copy -- no per-CTA gather.
struct PrepParams {
    const half* x;          // [M][ldx]
    int ldx, M, K, num_mats;
    const half* norm_w;     // or NULL
    float norm_eps;
    const uint16_t* perm[GEMV_MAX_MATS];
    half* xp[GEMV_MAX_MATS];
};
__global__ void __launch_bounds__(1024) tc_prep_kernel(const __grid_constant__ PrepParams P) {
    griddep_launch_dependents();
    griddep_wait();
    const int tok = blockIdx.x, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    __shared__ float part[32];
    float rstd = 1.f;
    const bool live = tok < P.M;
    const half* xr = P.x + (size_t)tok * P.ldx;
    if (live && P.norm_w) {            // cuda/rms_norm.cu:55-111: clamp, fp32 sum of squares, rsqrt(mean + eps)
        float sum = 0.f;
        for (int k = tid * 8; k < P.K; k += 1024 * 8) {
            const uint4 v4 = *reinterpret_cast<const uint4*>(xr + k);
            const half2* h2 = reinterpret_cast<const half2*>(&v4);
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                float f0 = fmaxf(-65504.f, fminf(__low2float(h2[i]), 65504.f));
                float f1 = fmaxf(-65504.f, fminf(__high2float(h2[i]), 65504.f));
                sum = fmaf(f0, f0, sum);
                sum = fmaf(f1, f1, sum);
            }
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, o);
        if (lane == 0) part[warp] = sum;
        __syncthreads();
        float t = (lane < 32) ? part[lane] : 0.f;
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) t += __shfl_xor_sync(0xffffffffu, t, o);
        rstd = rsqrtf(t * (1.0f / (float)P.K) + P.norm_eps);
    }
    for (int mi = 0; mi < P.num_mats; ++mi) {
        const uint16_t* perm = P.perm[mi];
        half* dst = P.xp[mi];
        for (int kp = tid; kp < P.K; kp += 1024) {
            half v = __float2half(0.f);
            if (live) {
                const int src = perm ? (int)__ldg(perm + kp) : kp;
                v = xr[src];
                if (P.norm_w) {
                    float xf = fmaxf(-65504.f, fminf(__half2float(v), 65504.f));
                    v = __float2half_rn(xf * __half2float(__ldg(P.norm_w + src)) * rstd);
                }
            }
            dst[(size_t)(kp >> 3) * 64 + tok * 8 + (kp & 7)] = v;
        }
    }
}

template <int MT>   // MT = 1 (decode) or 8 tokens per pass
__global__ void __launch_bounds__(TC_THREADS, 2) gemm_tc_kernel(const __grid_constant__ GemvParams P) {
    extern __shared__ __align__(1024) uint8_t smem[];
    const int tid = threadIdx.x, lane = tid & 31;
    const int warp = __shfl_sync(0xffffffffu, tid >> 5, 0);         // provably warp-uniform (UMMA operands live in uniform registers)
    const bool is_mma = warp >= TC_WARPS;                           // warps 8, 9: tensor-core issue for warpgroup 0, 1
    const int wg = is_mma ? warp - TC_WARPS : warp >> 2;            // the warpgroup this warp belongs to / serves
    const int rwg = is_mma ? 2 : wg;                                // role in the epilogue (issue warps take no part)
    const int wq = warp & 3, tidw = tid & 127;                      // TMEM lane quadrant, thread = weight column

    griddep_launch_dependents();
    TC_STAMP(0);

    // shared memory: [barriers][misc][WG1 totals][fp16 tile][ssq] | [2 activation rings] | [8 weight rings]
    const uint32_t smem0 = smem_addr(smem);
    const int stage_bytes = P.tc_stage_bytes, NS = P.tc_stages;
    const uint32_t wbar = smem0 + warp * TC_MAX_STAGES * 8;                                   // my weight stages
    const uint32_t abar = smem0 + (TC_WARPS * TC_MAX_STAGES + wg * TC_MAX_STAGES) * 8;          // my WG's activation stages
    const uint32_t barA = smem0 + (TC_WARPS * TC_MAX_STAGES + 2 * TC_MAX_STAGES + wg * 3) * 8;  // A buffer free (its MMAs retired)
    const uint32_t barD = smem0 + (TC_WARPS * TC_MAX_STAGES + 2 * TC_MAX_STAGES + 6 + wg * 2) * 8;   // accumulator complete
    const uint32_t barF = smem0 + (TC_WARPS * TC_MAX_STAGES + 2 * TC_MAX_STAGES + 10 + wg * 3) * 8;  // A buffer full (4 warps arrive)
    uint8_t* misc = smem + TC_SMEM_BARS;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(misc);
    int* flag_s = reinterpret_cast<int*>(misc + 64);
    float* rstd_s = reinterpret_cast<float*>(misc + 16);             // [8] deferred RMSNorm factors
    float* comb_s = reinterpret_cast<float*>(misc + TC_SMEM_MISC);  // [8][128] WG1 totals
    half* tile_s = reinterpret_cast<half*>(comb_s + GEMV_MTOK * 128);   // [8][128] fp16 outputs (RoPE partner exchange)
    float* ssq_s = reinterpret_cast<float*>(tile_s + GEMV_MTOK * 128);  // [4][8] per-warp sums of squares
    float* corr_s = ssq_s + 32 + wg * (TC_MAX_STAGES * 16);             // [stage][S1[8] | S0[8]] of my warpgroup's 4-bit groups
    const uint32_t act_ring = smem0 + P.tc_act_off + wg * NS * TC_ACT_STAGE;
    const int ring_off = P.tc_act_off + 2 * NS * TC_ACT_STAGE;
    uint8_t* ring_p = smem + ring_off + warp * NS * stage_bytes;
    const uint32_t ring = smem0 + ring_off + warp * NS * stage_bytes;
    const int M = P.M, KS = P.KS;

    // ---- one-time setup: barriers, TMEM ----
    if (lane == 0 && !is_mma) {
        for (int st = 0; st < NS; ++st) mbar_init(wbar + 8 * st, 1);
        if (wq == 0) {
            for (int st = 0; st < NS; ++st) mbar_init(abar + 8 * st, 1);
#pragma unroll
            for (int i = 0; i < 3; ++i) {
                mbar_init(barA + 8 * i, 1);
                mbar_init(barF + 8 * i, 4);
            }
            mbar_init(barD, 2);            // tcgen05.commit of the group's last chunk + the issue warp's own arrive (S1 / S0 written)
            mbar_init(barD + 8, 2);
        }
        mbar_fence_init();
    }
    __syncwarp();
    // (a warp's weight barriers are used by that warp alone: it may start fetching right away; the tensor-memory allocation
    //  and the CTA-wide barrier that publishes the shared barriers come after the first requests are in flight)
    bool setup_done = false;
    uint32_t tmem_base = 0, t_a = 0, t_d = 0;
    const uint32_t lane_sel = (uint32_t)(wq * 32) << 16;

    const unsigned U = (unsigned)P.total_units, G = gridDim.x;
    const int u0 = (int)((unsigned)blockIdx.x * U / G), u1 = (int)(((unsigned)blockIdx.x + 1u) * U / G);

    uint32_t wphase = 0, aphase = 0, phaseA = 0, phaseD = 0, phaseF = 0;     // parity bits per barrier
    TCP_DECL
    uint32_t ab = 0, a_uses = 0, dsel = 0;                        // A buffer / accumulator rotation of this WG (across segments)
    bool waited = false;
    auto after_wait = [&]() {       // first point where the previous kernel's output may be read
        if (P.ex.sumsq_in) {        // warp m: 1/rms of token m, strips summed in a fixed order
            float sum = 0.f;
            for (int sidx = lane; sidx < P.ex.sumsq_in_strips; sidx += 32) sum += __ldcg(P.ex.sumsq_in + sidx * 8 + warp);
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, o);
            if (lane == 0) rstd_s[warp] = rsqrtf(sum / (float)(KS * SLAB_K) + P.ex.sumsq_eps);
        }
    };

    int u = u0;
    while (u < u1) {
        int mi = 0;
        while (mi + 1 < P.num_mats && u >= P.mat[mi + 1].unit_begin) ++mi;
        const GemvMat& mt = P.mat[mi];
        const QMatView& w = mt.w;
        const int local = u - mt.unit_begin;
        const int strip = local / KS, ks_a = local - strip * KS;
        const int seg = min(KS - ks_a, u1 - u);
        // snap the slab range to group boundaries (every CTA applies the same rule, so ranges still tile the strip);
        // WG0 takes the first half of the range's groups, WG1 the second: two contiguous streams per CTA
        const int ks0 = tc_group_start(w, ks_a), ks1 = tc_group_start(w, ks_a + seg);
        const int ksm = tc_group_start(w, ks0 + ((ks1 - ks0 + 1) >> 1));
        const int my0 = wg ? ksm : ks0, my1 = wg ? ks1 : ksm;
        const int n_col = strip * 128 + tidw;                                  // this thread's output column
        const bool col_live = n_col < w.N;
        const uint8_t* gsrc = reinterpret_cast<const uint8_t*>(w.packed) + (size_t)strip * w.strip_bytes + (size_t)wq * w.blk_stream_bytes;
        const uint8_t* asrc = reinterpret_cast<const uint8_t*>(mt.xp);
        const int nreg = w.num_regions;
        // per-matrix fields used in the loops below, pinned in registers (indexed kernel-parameter reads are slow and the
        // compiler would otherwise re-read them every group)
        const bool gptq = w.is_gptq != 0;
        const uint32_t* sc_w = (gptq ? w.qzeros : w.q_scale) + (n_col >> 3);     // my column's nibble word of group 0
        const half* sc_h = gptq ? w.gptq_scales + n_col : w.q_scale_max;
        int n8 = w.N >> 3;
        asm volatile("" : "+l"(sc_w));
        asm volatile("" : "+l"(sc_h));
        asm volatile("" : "+r"(n8));
        const int nib_sh = (n_col & 7) * 4;

        // Group cursors: position + the current region's parameters in registers, so that stepping to the next group
        // is a handful of integer ops; the kernel-parameter region table is only read when a region boundary is crossed.
        //   C: the group being unpacked;   F: the next group to request (weights + activations), NS groups ahead
        int c_ks = my0, c_r = tc_region_of(w, min(my0, KS - 1));
        int c_bits = w.reg[c_r].bits, c_spg = 1 << w.reg[c_r].spg_log2, c_end = tc_region_end(w, c_r);
        int c_grp = w.reg[c_r].group_base + ((my0 - w.reg[c_r].ks_begin) >> w.reg[c_r].spg_log2);
        int f_ks = c_ks, f_r = c_r, f_bits = c_bits, f_spg = c_spg, f_end = c_end;
        uint32_t f_off = w.reg[c_r].off_base + (uint32_t)(my0 - w.reg[c_r].ks_begin) * (uint32_t)block_bytes(c_bits);
        int f_stage = 0, cstage = 0;

        // request group F's weights (every warp: its own block) and, optionally, its activations (the WG's first warp)
        auto issue = [&](bool weights, bool acts) {
            const int ns = min(f_spg, f_end - f_ks);
            const uint32_t wbytes = (uint32_t)ns * (uint32_t)block_bytes(f_bits);
            if (elect_one()) {
                if (weights) {
                    mbar_arrive_expect_tx(wbar + 8 * f_stage, wbytes);
                    bulk_copy_g2s(ring + f_stage * stage_bytes, gsrc + f_off, wbytes, wbar + 8 * f_stage);
                }
                if (acts && wq == 0) {
                    mbar_arrive_expect_tx(abar + 8 * f_stage, (uint32_t)ns * SLAB_K * 16);
                    bulk_copy_g2s(act_ring + f_stage * TC_ACT_STAGE, asrc + (size_t)f_ks * SLAB_K * 16, (uint32_t)ns * SLAB_K * 16,
                                  abar + 8 * f_stage);
                }
            }
            f_ks += ns;
            f_off += wbytes;
            f_stage = (f_stage + 1 == NS) ? 0 : f_stage + 1;
            if (f_ks >= f_end && f_r + 1 < nreg) {
                ++f_r;
                f_bits = w.reg[f_r].bits;
                f_spg = 1 << w.reg[f_r].spg_log2;
                f_end = tc_region_end(w, f_r);
                f_off = w.reg[f_r].off_base;
            }
        };
        // prologue: the first NS groups' weights (they never depend on a previous kernel); their activations follow after
        // griddepcontrol.wait, requested by re-walking the same groups with a scratch copy of the cursor
        int primed = 0;
#pragma unroll 1
        for (; !is_mma && primed < NS && f_ks < my1; ++primed) issue(true, false);
        if (!setup_done) {
            setup_done = true;
            if (warp == 0) tmem_alloc(smem_addr(tmem_slot), TC_TMEM_COLS);
            tc_fence_before();
            __syncthreads();
            tc_fence_after();
            tmem_base = *tmem_slot;
            t_a = tmem_base + wg * TC_COLS_PER_WG;          // A: 3 x 32 columns, D: 2 x 16
            t_d = t_a + TC_A_BUFS * 32;
        }
        TC_STAMP(1);

        if (is_mma) {
            // ---- tensor-core issue warp of warpgroup wg: for every chunk wait until the four unpack warps have filled the
            //      A buffer, issue its MMAs (2 per slab, K = 16), commit to "A buffer free" (+ "accumulator complete" at the
            //      group's last chunk).  It never blocks the unpack warps: they only meet it through mbarriers.
            int m_ks = my0, m_r = c_r, m_spg = c_spg, m_end = c_end, m_bits = c_bits, mstage = 0;
            while (m_ks < my1) {
                const int ns = min(m_spg, m_end - m_ks);
                const int nchunks = (ns + 1) >> 1;
                TCP_BEGIN;
                mbar_wait(abar + 8 * mstage, (aphase >> mstage) & 1u);       // the group's activations have landed
                aphase ^= 1u << mstage;
                TCP_END(0);
                const uint8_t* actp = smem + P.tc_act_off + (wg * NS + mstage) * TC_ACT_STAGE;
#pragma unroll 1
                for (int ch = 0; ch < nchunks; ++ch) {
                    const int nsl = min(2, ns - 2 * ch);
                    TCP_BEGIN;
                    mbar_wait(barF + 8 * ab, (phaseF >> ab) & 1u);
                    phaseF ^= 1u << ab;
                    tc_fence_after();
                    TCP_END(0);
                    if (ch == 0) {
                        // (all four unpack warps are past the read-back of the group that last used this stage's S1 / S0)
                        if (m_bits == 4) {
                            // lane = (token t, k-quarter kq): sums over the 8-k core-matrix rows 4*kq .. 4*kq+3 of every slab
                            const int t = lane & 7, kq = lane >> 3;
                            float s1 = 0.f, s0 = 0.f;
                            for (int row = kq; row < ns * 4; row += 4) {
                                const uint4 v = *reinterpret_cast<const uint4*>(actp + row * 128 + t * 16);
                                const float2 f0 = __half22float2(*reinterpret_cast<const half2*>(&v.x)), f1 = __half22float2(*reinterpret_cast<const half2*>(&v.y));
                                const float2 f2 = __half22float2(*reinterpret_cast<const half2*>(&v.z)), f3 = __half22float2(*reinterpret_cast<const half2*>(&v.w));
                                const float e = (f0.x + f0.y) + (f2.x + f2.y), o = (f1.x + f1.y) + (f3.x + f3.y);   // even / odd pair slots
                                s1 = fmaf(1024.f, e, fmaf(64.f, o, s1));
                                s0 += e + o;
                            }
                            s1 += __shfl_xor_sync(0xffffffffu, s1, 8);
                            s0 += __shfl_xor_sync(0xffffffffu, s0, 8);
                            s1 += __shfl_xor_sync(0xffffffffu, s1, 16);
                            s0 += __shfl_xor_sync(0xffffffffu, s0, 16);
                            if (lane < 8) {
                                corr_s[mstage * 16 + lane] = s1;
                                corr_s[mstage * 16 + 8 + lane] = s0;
                            }
                        }
                        __syncwarp();
                        if (elect_one()) mbar_arrive(barD + 8 * dsel);
                    }
                    TCP_END(1);
                    // K-step j: A advances 8 TMEM columns, B advances 2 core matrices = 256 B = 16 descriptor units
                    const uint64_t bd0 = make_b_desc(act_ring + mstage * TC_ACT_STAGE + (uint32_t)ch * 1024u);
                    const uint32_t td = t_d + dsel * 16, ta = t_a + ab * 32;
                    if (elect_one()) {
                        if (nsl == 2) {
#pragma unroll
                            for (int j = 0; j < 4; ++j) umma_ts(td, ta + j * 8, bd0 + (uint64_t)(j * 16), TC_IDESC, (ch | j) ? 1u : 0u);
                        } else {
#pragma unroll
                            for (int j = 0; j < 2; ++j) umma_ts(td, ta + j * 8, bd0 + (uint64_t)(j * 16), TC_IDESC, (ch | j) ? 1u : 0u);
                        }
#ifdef EXL2B_TC_PROFILE
                    }
                    __syncwarp();
                    TCP_END(2);
                    if (elect_one()) {
#endif
                        umma_commit(barA + 8 * ab);
                        if (ch == nchunks - 1) umma_commit(barD + 8 * dsel);
                    }
                    __syncwarp();
                    TCP_END(3);
                    ab = (ab == 2u) ? 0u : ab + 1u;
                }
                dsel ^= 1u;
                mstage = (mstage + 1 == NS) ? 0 : mstage + 1;
                m_ks += ns;
                if (m_ks >= m_end && m_r + 1 < nreg) {
                    ++m_r;
                    m_spg = 1 << w.reg[m_r].spg_log2;
                    m_end = tc_region_end(w, m_r);
                    m_bits = w.reg[m_r].bits;
                }
            }
        }


        // everything above depended only on the weights; from here on the previous kernel's output is needed
        if (!waited && !is_mma) {
            griddep_wait();
            waited = true;
            after_wait();
            TC_STAMP(2);
        }
        if (!is_mma && wq == 0) {            // the activations of the groups primed above (same walk, scratch cursor)
            int t_ks = my0, t_r = c_r, t_spg = c_spg, t_end = c_end;
            for (int st = 0; st < primed; ++st) {
                const int tn = min(t_spg, t_end - t_ks);
                if (elect_one()) {
                    mbar_arrive_expect_tx(abar + 8 * st, (uint32_t)tn * SLAB_K * 16);
                    bulk_copy_g2s(act_ring + st * TC_ACT_STAGE, asrc + (size_t)t_ks * SLAB_K * 16, (uint32_t)tn * SLAB_K * 16, abar + 8 * st);
                }
                t_ks += tn;
                if (t_ks >= t_end && t_r + 1 < nreg) {
                    ++t_r;
                    t_spg = 1 << w.reg[t_r].spg_log2;
                    t_end = tc_region_end(w, t_r);
                }
            }
        }

        // ---- the unpack pipeline of a warp.  Per chunk (<= 2 slabs = 64 k): wait for a free A buffer -> unpack ->
        //      tcgen05.st -> arrive on "A full"; the issue warp does the rest.  The accumulator of group g is read back
        //      (tcgen05.ld, scaled, added) only after group g+1 has been unpacked, so tensor core and unpack overlap.
        float tot[MT];
#pragma unroll
        for (int m = 0; m < MT; ++m) tot[m] = 0.f;
        bool pending = false;
        // State of a group between its unpack and its read-back.  Two instances used alternately (the loop below is unrolled
        // by two): a scale loaded at the top of group g is first touched when g is read back, at the END of group g+1 -- no
        // register move in between that would make the in-order warp wait for the load.
        struct GroupState {
            uint32_t word;       // raw q_scale word (EXL2)
            half h;              // q_scale_max[group] (EXL2) / scale of (group, column) (GPTQ)
            int z;               // zero point when unpacked in two-offset form (EXL2: 8, GPTQ: z + 1), -1: exact unpack
            uint32_t d;          // accumulator index
            int stage;           // pipeline stage (where the issue warp left S1 / S0)
        };
        GroupState ga = {0u, __float2half(0.f), -1, 0u, 0}, gb = ga;
        uint32_t zw_next = 0;            // GPTQ: qzeros word of the NEXT group, fetched a group ahead
        if (!is_mma && gptq && col_live && c_ks < my1) zw_next = __ldg(sc_w + (size_t)c_grp * n8);

        auto drain = [&](const GroupState& g) {       // accumulator of group g -> running totals; its stage gets the next request
            const uint32_t d = g.d;
            mbar_wait(barD + 8 * d, (phaseD >> d) & 1u);
            phaseD ^= 1u << d;
            tc_fence_after();
            TCP_END(2);
            float pend_scale = 0.f;
            if (col_live) {
                if (!gptq) {
                    const int nib = (int)((g.word >> nib_sh) & 15u);
                    pend_scale = __half2float(__hmul(__int2half_rn((nib + 1) * (nib + 1)), g.h));    // qdq_util.cuh:24-30
                } else {
                    pend_scale = __half2float(g.h);
                }
            }
            float dd[MT];
            if constexpr (MT == 1) tmem_ld1(t_d + d * 16 + lane_sel, dd[0]);
            else tmem_ld8(t_d + d * 16 + lane_sel, dd);
            if (g.z >= 0) {              // remove the unpack offsets and the zero point
                const float zf = (float)g.z;
                const float* cs = corr_s + g.stage * 16;
                tmem_wait_ld();
#pragma unroll
                for (int m = 0; m < MT; ++m) tot[m] = fmaf(pend_scale, dd[m] - fmaf(zf, cs[8 + m], cs[m]), tot[m]);
            } else {
                tmem_wait_ld();
#pragma unroll
                for (int m = 0; m < MT; ++m) tot[m] = fmaf(pend_scale, dd[m], tot[m]);
            }
            tc_fence_before();
            if (f_ks < my1) issue(true, true);       // the drained group's MMAs have retired: both of its stages are free
        };

        auto group_step = [&](GroupState& cur, const GroupState& prev) {
            const int bits = c_bits;
            const int ns = min(c_spg, c_end - c_ks);
            const int grp = c_grp;

            // (a) this group's scale for my column: requested now, decoded when the group is read back
            cur.word = 0u;
            cur.h = __float2half(0.f);
            cur.z = bits == 4 ? 8 : -1;
            if (col_live) {
                if (!gptq) {
                    cur.word = __ldg(sc_w + (size_t)grp * n8);
                    cur.h = __ldg(sc_h + grp);
                } else {
                    cur.h = __ldg(sc_h + (size_t)grp * (n8 * 8));
                    cur.z = (int)((zw_next >> nib_sh) & 15u) + 1;                                   // q_gemm_kernel_gptq.cuh:167-172
                    if (c_ks + ns < my1) zw_next = __ldg(sc_w + (size_t)(grp + 1) * n8);           // GPTQ: one region
                }
            }
            cur.d = dsel;
            cur.stage = cstage;

            // (b) my block's slabs of this group have landed
            TCP_BEGIN;
            mbar_wait(wbar + 8 * cstage, (wphase >> cstage) & 1u);
            wphase ^= 1u << cstage;
            TCP_END(0);
            const uint8_t* sp = ring_p + cstage * stage_bytes;
            const int nchunks = (ns + 1) >> 1;
            auto run_chunks = [&](auto tag) {
                constexpr int B = decltype(tag)::value;
#pragma unroll 1
                for (int ch = 0; ch < nchunks; ++ch) {
                    const int nsl = min(2, ns - 2 * ch);
                    TCP_BEGIN;
                    if (a_uses >= 3u) {              // the MMAs that read this A buffer three chunks ago have retired
                        mbar_wait(barA + 8 * ab, (phaseA >> ab) & 1u);
                        phaseA ^= 1u << ab;
                        tc_fence_after();
                    }
                    TCP_END(0);
                    tc_dequant_chunk<B>(sp, 2 * ch, nsl, lane, t_a + ab * 32 + lane_sel);
                    tmem_wait_st();
                    tc_fence_before();
                    __syncwarp();
                    if (elect_one()) mbar_arrive(barF + 8 * ab);     // 1 of 4: this warp's 32 rows of the A buffer are in place
                    ab = (ab == 2u) ? 0u : ab + 1u;
                    ++a_uses;
                    TCP_END(1);
                }
            };
            switch (bits) {
                case 4: run_chunks(std::integral_constant<int, 4>{}); break;
                case 5: run_chunks(std::integral_constant<int, 5>{}); break;
                case 3: run_chunks(std::integral_constant<int, 3>{}); break;
                case 6: run_chunks(std::integral_constant<int, 6>{}); break;
                case 2: run_chunks(std::integral_constant<int, 2>{}); break;
                default: run_chunks(std::integral_constant<int, 8>{}); break;
            }
            // (c) read back the PREVIOUS group while this one's MMAs run (and re-arm its stage), then step the cursor
            TCP_BEGIN;
            if (pending) drain(prev);
            TCP_END(3);
            pending = true;
            dsel ^= 1u;
            cstage = (cstage + 1 == NS) ? 0 : cstage + 1;
            c_ks += ns;
            ++c_grp;
            if (c_ks >= c_end && c_r + 1 < nreg) {
                ++c_r;
                c_bits = w.reg[c_r].bits;
                c_spg = 1 << w.reg[c_r].spg_log2;
                c_end = tc_region_end(w, c_r);
                c_grp = w.reg[c_r].group_base;
            }
        };
        bool last_a = false;
        while (!is_mma && c_ks < my1) {
            group_step(ga, gb);
            last_a = true;
            if (c_ks >= my1) break;
            group_step(gb, ga);
            last_a = false;
        }
        if (pending) {
            if (last_a) drain(ga);
            else drain(gb);
        }
        if (!waited && !is_mma) {        // a CTA whose snapped range is empty still takes part in the fix-up below
            griddep_wait();
            waited = true;
            after_wait();
        }
        TC_STAMP(4);

        // the finishing CTA's scatter needs, per consumer, this column's destination row and RMSNorm weight: request them now, so
        // that the loads ride under the combine / split-K hand-off below instead of sitting on the tail of the launch
        int skp[GEMV_MAX_MATS];
        half ssc[GEMV_MAX_MATS];
        half resid0 = __float2half(0.f);            // decode (one row): the residual element this column adds to, same idea
        if constexpr (MT == 1) {
            if (rwg == 0 && P.epilogue == EPI_STORE && !mt.clear && n_col < w.N) resid0 = mt.c[n_col];
        }
        {
            const bool col_any = n_col < ((P.epilogue != EPI_STORE) ? P.mat[0].w.N : w.N);
#pragma unroll
            for (int t = 0; t < GEMV_MAX_MATS; ++t) {
                skp[t] = n_col;
                ssc[t] = __float2half(1.f);
                if (rwg == 0 && t < P.ex.num_scat && col_any) {
                    if (P.ex.scat[t].invperm) skp[t] = (int)__ldg(P.ex.scat[t].invperm + n_col);
                    if (P.ex.scat[t].scale) ssc[t] = __ldg(P.ex.scat[t].scale + n_col);
                }
            }
        }

        // ---- combine the two warpgroups, then the split-K / epilogue logic of the mma.sync kernel ----
        __syncthreads();
        if (rwg == 1) {
#pragma unroll
            for (int m = 0; m < MT; ++m) comb_s[m * 128 + tidw] = tot[m];
        }
        __syncthreads();

        const int gs = mt.strip_begin + strip;
        const unsigned sb = (unsigned)mt.unit_begin + (unsigned)strip * KS;
        const int first_cta = tc_cta_of_unit(sb, G, U), last_cta = tc_cta_of_unit(sb + KS - 1, G, U);
        const int nc = last_cta - first_cta + 1, jc = (int)blockIdx.x - first_cta;
        const bool paired = P.epilogue != EPI_STORE;
        const bool has_rstd = P.ex.sumsq_in != nullptr;
        if (rwg == 0) {
#pragma unroll
            for (int m = 0; m < MT; ++m) tot[m] += comb_s[m * 128 + tidw];
        }
        // fin / fin2: the strip's complete fp32 sums (fin2 = the up projection of a gate/up pair), held by the finishing CTA
        bool finisher = false;
        float fin[MT], fin2[MT];
#pragma unroll
        for (int m = 0; m < MT; ++m) { fin[m] = tot[m]; fin2[m] = 0.f; }
        if (nc == 1 && !paired) {
            finisher = true;
        } else {
            float* wsp = P.ws + ((size_t)gs * P.maxc + jc) * TC_RED_FLOATS;
            if (rwg == 0) {
#pragma unroll
                for (int m = 0; m < MT; ++m) __stcg(wsp + m * 128 + tidw, tot[m]);
            }
            __syncthreads();             // every partial of this CTA is written (CTA-scope happens-before to thread 0)
            int expected = nc, cidx = gs;
            if (paired) {
                const GemvMat& other = P.mat[1 - mi];
                const unsigned ob = (unsigned)other.unit_begin + (unsigned)strip * KS;
                expected += tc_cta_of_unit(ob + KS - 1, G, U) - tc_cta_of_unit(ob, G, U) + 1;
                cidx = P.mat[0].strip_begin + strip;
            }
            if (tid == 0) {              // release our partials / acquire everybody else's: one acq_rel RMW at GPU scope
                unsigned int old;
                asm volatile("atom.add.acq_rel.gpu.global.u32 %0, [%1], %2;" : "=r"(old) : "l"(P.counters + cidx), "r"(1u) : "memory");
                *flag_s = (old == (unsigned int)(expected - 1)) ? 1 : 0;
            }
            __syncthreads();
            if (*flag_s) {
                finisher = true;
                if (rwg == 0) {
                    if (!paired) {
                        const float* base = P.ws + (size_t)gs * P.maxc * TC_RED_FLOATS;
#pragma unroll
                        for (int m = 0; m < MT; ++m) {
                            float v = 0.f;
                            for (int j = 0; j < nc; ++j) v += __ldcg(base + (size_t)j * TC_RED_FLOATS + m * 128 + tidw);
                            fin[m] = v;
                        }
                    } else {
                        const GemvMat& mg = P.mat[0];
                        const GemvMat& mu = P.mat[1];
                        const unsigned gb = (unsigned)mg.unit_begin + (unsigned)strip * KS;
                        const unsigned ub = (unsigned)mu.unit_begin + (unsigned)strip * KS;
                        const int ncg = tc_cta_of_unit(gb + KS - 1, G, U) - tc_cta_of_unit(gb, G, U) + 1;
                        const int ncu = tc_cta_of_unit(ub + KS - 1, G, U) - tc_cta_of_unit(ub, G, U) + 1;
                        const float* bg = P.ws + (size_t)(mg.strip_begin + strip) * P.maxc * TC_RED_FLOATS;
                        const float* bu = P.ws + (size_t)(mu.strip_begin + strip) * P.maxc * TC_RED_FLOATS;
#pragma unroll
                        for (int m = 0; m < MT; ++m) {
                            float vg = 0.f, vu = 0.f;
                            for (int j = 0; j < ncg; ++j) vg += __ldcg(bg + (size_t)j * TC_RED_FLOATS + m * 128 + tidw);
                            for (int j = 0; j < ncu; ++j) vu += __ldcg(bu + (size_t)j * TC_RED_FLOATS + m * 128 + tidw);
                            fin[m] = vg;
                            fin2[m] = vu;
                        }
                    }
                }
                if (tid == 0) P.counters[cidx] = 0u;
            }
        }

        // ---- the finishing CTA's first warpgroup turns the sums into outputs (thread = column) ----
        if (finisher && rwg == 0) {
            const GemvMat& mo = paired ? P.mat[0] : mt;         // where the result goes
            const bool col_ok = n_col < mo.w.N;
            half hv[MT];
#pragma unroll
            for (int m = 0; m < MT; ++m) {
                const float rs = has_rstd ? rstd_s[m] : 1.f;
                if (paired) {
                    float vg = fin[m] * rs, vu = fin2[m] * rs;
                    if (col_ok && P.mat[0].w.bias) vg += __half2float(P.mat[0].w.bias[n_col]);
                    if (col_ok && P.mat[1].w.bias) vu += __half2float(P.mat[1].w.bias[n_col]);
                    const half hg = __float2half_rn(vg), hu = __float2half_rn(vu);           // q_mlp.cu:187-196 roundings
                    const half act = (P.epilogue == EPI_GELU_MUL) ? tc_gelu_h(hg) : tc_silu_h(hg);
                    hv[m] = __hmul(act, hu);
                } else {
                    float v = fin[m] * rs;
                    if (col_ok && m < M) {
                        if (w.bias) v += __half2float(w.bias[n_col]);
                        if (!mt.clear) v += __half2float(MT == 1 ? resid0 : mt.c[(size_t)m * mt.ldc + n_col]);
                    }
                    hv[m] = __float2half_rn(v);
                }
            }
            if ((P.ex.rope.mask >> mi) & 1u) {          // RoPE on the fp16 values, partner through shared memory
                const RopeFuse& R = P.ex.rope;
#pragma unroll
                for (int m = 0; m < MT; ++m) tile_s[m * 128 + tidw] = hv[m];
                bar_sync(3, 128);
                const int d = tidw % R.head_dim, S = R.sincos_size, hd2 = S >> 1;
#pragma unroll
                for (int m = 0; m < MT; ++m) {
                    if (m < M && d < S) {
                        const int row = P.row0 + m, bb = row / R.q_len, tt = row - bb * R.q_len;
                        int base = R.past_len;
                        if (base == -1) base = max(R.past_lens[bb], 0);
                        else if (R.past_lens) base += R.past_lens[bb];
                        const size_t sr = (size_t)max(base + tt, 0) * S;
                        if (R.neox) {
                            if (d < hd2) {
                                const half c = R.cos[sr + d], sn = R.sin[sr + d];
                                hv[m] = __hfma(hv[m], c, __hmul(tile_s[m * 128 + tidw + hd2], __hneg(sn)));
                            } else {
                                const half c = R.cos[sr + d - hd2], sn = R.sin[sr + d - hd2];
                                hv[m] = __hfma(hv[m], c, __hmul(tile_s[m * 128 + tidw - hd2], sn));
                            }
                        } else {
                            const half c = R.cos[sr + d], sn = R.sin[sr + d];
                            if ((d & 1) == 0) hv[m] = __hfma(tile_s[m * 128 + tidw + 1], __hneg(sn), __hmul(hv[m], c));
                            else hv[m] = __hfma(tile_s[m * 128 + tidw - 1], sn, __hmul(hv[m], c));
                        }
                    }
                }
            }
            float ssq[MT];
#pragma unroll
            for (int m = 0; m < MT; ++m) {
                ssq[m] = 0.f;
                if (col_ok && m < M) {
                    mo.c[(size_t)m * mo.ldc + n_col] = hv[m];
                    const float f = fmaxf(-65504.f, fminf(__half2float(hv[m]), 65504.f));
                    ssq[m] = f * f;
#pragma unroll
                    for (int t = 0; t < GEMV_MAX_MATS; ++t) {
                        if (t < P.ex.num_scat) {
                            const ScatterTarget& T = P.ex.scat[t];
                            const int kp = skp[t];
                            const half o = T.scale ? __float2half_rn(f * __half2float(ssc[t])) : hv[m];
                            T.xp[(size_t)(kp >> 3) * 64 + m * 8 + (kp & 7)] = o;
                        }
                    }
                }
            }
            if (P.ex.sumsq_out) {        // per-strip sum of squares of the stored rows, fixed reduction order
#pragma unroll
                for (int m = 0; m < MT; ++m) {
#pragma unroll
                    for (int o = 16; o > 0; o >>= 1) ssq[m] += __shfl_xor_sync(0xffffffffu, ssq[m], o);
                    if (lane == 0) ssq_s[wq * 8 + m] = ssq[m];
                }
                bar_sync(3, 128);
                if (tidw < 8) {
                    const float t4 = (ssq_s[tidw] + ssq_s[8 + tidw]) + (ssq_s[16 + tidw] + ssq_s[24 + tidw]);
                    P.ex.sumsq_out[(size_t)strip * 8 + tidw] = (tidw < MT) ? t4 : 0.f;
                }
            }
        }
        __syncthreads();
        TC_STAMP(5);
        TCP_FLUSH;
        u += seg;
    }

    tc_fence_before();
    __syncthreads();
    if (warp == 0) tmem_dealloc(tmem_base, TC_TMEM_COLS);
    if (P.dbg && tid == 0) atomicMax(P.dbg + 7, tc_gtimer());
}

// ---- host launcher -----------------------------------------------------------------------------------------------------

int gemv_workspace(int device, cudaStream_t stream, float** ws, unsigned int** counters, size_t* ws_bytes, int* n_counters, half** xp,
                   size_t xp_bytes);

constexpr size_t TC_XP_BYTES_PER_MAT = (size_t)65536 * 16;      // K <= 65536, 16 B per k (8 token slots)

int gemm_tc_launch(int device, cudaStream_t stream, GemvMat* mats, int nm, int M, const half* norm_w, float norm_eps, int epilogue,
                   const GemvExtras* ex) {
    EXL2B_REQUIRE(nm >= 1 && nm <= GEMV_MAX_MATS, "bad matrix count %d", nm);
    if (M <= 0) return 0;
    float* ws = nullptr;
    unsigned int* counters = nullptr;
    size_t ws_bytes = 0;
    int n_counters = 0;
    half* xp_scratch = nullptr;
    int rc = gemv_workspace(device, stream, &ws, &counters, &ws_bytes, &n_counters, &xp_scratch, TC_XP_BYTES_PER_MAT * GEMV_MAX_MATS);
    if (rc) return rc;
    static std::atomic<bool> attr_set[64];
    if (!attr_set[device].load()) {
        EXL2B_CUDA(cudaFuncSetAttribute(gemm_tc_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
        EXL2B_CUDA(cudaFuncSetAttribute(gemm_tc_kernel<8>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
        attr_set[device].store(true);
    }

    GemvParams P = {};
    P.num_mats = nm;
    P.KS = mats[0].w.KS;
    const bool prepared = ex && ex->prepared;
    if (ex) {
        P.ex = *ex;
        EXL2B_REQUIRE(M <= GEMV_MTOK, "fused epilogue extras need a single pass (rows %d > %d)", M, GEMV_MTOK);
        if (ex->rope.mask) EXL2B_REQUIRE(ex->rope.head_dim > 0 && 128 % ex->rope.head_dim == 0 && ex->rope.sincos_size <= ex->rope.head_dim,
                                         "fused RoPE needs head_dim to divide 128");
        if (ex->sumsq_in) P.ex.sumsq_eps = norm_eps;
    }
    long long units = 0;
    int strips = 0;
    for (int i = 0; i < nm; ++i) {
        EXL2B_REQUIRE(mats[i].w.layout == LAYOUT_TC, "matrix is not in the tcgen05 layout");
        EXL2B_REQUIRE(mats[i].w.KS == P.KS, "fused matrices must share K");
        EXL2B_REQUIRE(prepared || (mats[i].x == mats[0].x && mats[i].ldx == mats[0].ldx), "fused matrices must share their input");
        P.mat[i] = mats[i];
        P.mat[i].unit_begin = (int)units;
        P.mat[i].strip_begin = strips;
        if (prepared) EXL2B_REQUIRE(mats[i].xp, "prepared launch without an activation buffer");
        else P.mat[i].xp = xp_scratch + (size_t)i * (TC_XP_BYTES_PER_MAT / sizeof(half));
        units += (long long)mats[i].w.strips * P.KS;
        strips += mats[i].w.strips;
    }
    if (norm_w && !prepared) EXL2B_REQUIRE(mats[0].w.K % 8 == 0 && mats[0].ldx % 8 == 0, "fused RMSNorm needs K and the row stride to be multiples of 8");
    if (epilogue != EPI_STORE) EXL2B_REQUIRE(nm == 2 && mats[0].w.N == mats[1].w.N, "gate/up epilogue needs two matrices of equal width");
    P.norm_w = norm_w;
    P.norm_eps = norm_eps;
    P.epilogue = epilogue;
    P.ws = ws;
    P.counters = counters;

    // Grid: every strip cut into S equal K-ranges (one segment per CTA) or plain stream-K over all resident slots --
    // whichever has the cheaper slowest CTA under  cost = segments * F + slabs  (F = per-segment fixed cost in slabs).
    extern int g_tc_ctas_per_sm;
    const int sms = device_sm_count(device);
    const long long slots = (long long)sms * g_tc_ctas_per_sm;
    const long long F = 16;
    long long grid_ll = std::min(slots, units);
    {
        const long long L = (units + grid_ll - 1) / grid_ll;
        const long long cost_stream = ((L + P.KS - 1) / P.KS + 1) * F + L;
        if (strips <= slots) {
            // S need not divide KS: CTA i covers units [i*U/G, (i+1)*U/G), every S-th boundary is a strip boundary, the others
            // are snapped to quantisation-group starts by the kernel -- still exactly one segment per CTA
            const int S = (int)std::min<long long>(slots / strips, std::max(1, P.KS / 8));
            const long long cost_aligned = F + (P.KS + S - 1) / S;
            if (cost_aligned <= cost_stream) grid_ll = (long long)strips * S;
        }
    }
    const int grid = (int)std::max(1ll, grid_ll);
    EXL2B_REQUIRE((units + 1) * grid < (1ll << 31), "problem too large for 32-bit unit arithmetic");
    P.total_units = (int)units;
    P.maxc = (int)(((long long)P.KS * grid) / units) + 2;
    EXL2B_REQUIRE(strips <= n_counters, "too many strips for the counter array");
    EXL2B_REQUIRE((size_t)strips * P.maxc * TC_RED_FLOATS * sizeof(float) <= ws_bytes, "split-K workspace too small");

    // stage = the largest group of any matrix of the launch, per 32-column block
    int stage_bytes = 0;
    for (int i = 0; i < nm; ++i)
        for (int r = 0; r < mats[i].w.num_regions; ++r)
            stage_bytes = std::max(stage_bytes, (1 << mats[i].w.reg[r].spg_log2) * block_bytes(mats[i].w.reg[r].bits));
    EXL2B_REQUIRE(stage_bytes > 0 && stage_bytes <= 4096, "quantisation groups above 128 rows are not supported by the tcgen05 kernel");
    P.tc_stage_bytes = stage_bytes;
    const int header = ((TC_SMEM_BARS + TC_SMEM_MISC + GEMV_MTOK * 128 * 4 + GEMV_MTOK * 128 * 2 + 128 + 2 * TC_MAX_STAGES * 64 + 1023) / 1024) * 1024;
    P.tc_act_off = header;
    // as many stages (<= 4) as leave room for the intended number of CTAs per SM (227 KB of shared memory, 1 KB reserved per CTA)
    const size_t smem_budget = (size_t)(227 * 1024) / (size_t)std::max(1, g_tc_ctas_per_sm) - 1024;
    int stages = TC_MAX_STAGES;
    auto smem_for = [&](int st) { return (size_t)header + (size_t)st * (2 * TC_ACT_STAGE + (size_t)TC_WARPS * stage_bytes); };
    while (stages > 2 && smem_for(stages) > smem_budget) --stages;
    P.tc_stages = stages;
    P.tc_act_bytes = 2 * stages * TC_ACT_STAGE;
    const size_t smem_total = smem_for(stages);
    EXL2B_REQUIRE(smem_total <= 200 * 1024, "shared memory budget exceeded (%zu bytes)", smem_total);
    EXL2B_REQUIRE((size_t)mats[0].w.K * 16 <= TC_XP_BYTES_PER_MAT, "K too large for the activation scratch");
    extern unsigned long long* g_dbg;
    extern int g_dbg_cta, g_dbg_slot;
    P.dbg_cta = g_dbg_cta;

    PrepParams Q = {};
    Q.ldx = mats[0].ldx;
    Q.K = mats[0].w.K;
    Q.num_mats = nm;
    Q.norm_w = norm_w;
    Q.norm_eps = norm_eps;
    for (int i = 0; i < nm; ++i) {
        Q.perm[i] = mats[i].w.perm;
        Q.xp[i] = const_cast<half*>(P.mat[i].xp);
    }
    for (int m0 = 0; m0 < M; m0 += GEMV_MTOK) {
        P.M = std::min(GEMV_MTOK, M - m0);
        P.dbg = g_dbg ? g_dbg + 32 * (g_dbg_slot++ % 64) : nullptr;
        for (int i = 0; i < nm; ++i) {
            if (!prepared) P.mat[i].x = mats[i].x + (size_t)m0 * mats[i].ldx;
            P.mat[i].c = mats[i].c + (size_t)m0 * mats[i].ldc;
        }
        P.row0 = m0;
        if (!prepared) {
            Q.x = mats[0].x + (size_t)m0 * mats[0].ldx;
            Q.M = P.M;
            EXL2B_CUDA(launch_pdl_f("tc", tc_prep_kernel, dim3(GEMV_MTOK), dim3(1024), 0, stream, Q));
        }
        if (P.M == 1) {
            EXL2B_CUDA(launch_pdl_f("tc", gemm_tc_kernel<1>, dim3(grid), dim3(TC_THREADS), smem_total, stream, P));
        } else {
            EXL2B_CUDA(launch_pdl_f("tc", gemm_tc_kernel<8>, dim3(grid), dim3(TC_THREADS), smem_total, stream, P));
        }
    }
    return 0;
}

}  // namespace exl2b
