// gemv_i8_kernel -- the batch-1 (decode) dequant-GEMV of libexl2b200, HBM-bound by design.
// Replaces, for one input row, gemm_half_q_half_kernel / gemm_half_q_half_gptq_kernel (exllamav2_ext/cuda/q_gemm_kernel.cuh:
// 140-565, q_gemm_kernel_gptq.cuh:60-225) together with the kernels the reference runs in front of them (rms_norm_kernel,
// cuda/rms_norm.cu:55-143; act_mul_kernel, cuda/q_mlp_activation.cuh:54-100), which here are the kernel's PROLOGUE.
//
// Why integers: measured on B200 (tools/ubench/dp4a.cu, profiles/r02_ubench_dp4a_pdlchain.txt) the reference's recipe
// (unpack to fp16, HFMA2) tops out at 38-44 4-bit weights/clk/SM while HBM delivers 45; the same loop on the integer
// dot-product instruction (IDP.4A) runs at 75-98.  So the row is quantised ONCE per launch to 16-bit integers per 128-k
// block with a power-of-two scale (signed high byte plane + unsigned low byte plane; values within 16x of the block
// maximum are exact, the rest carry |error| <= 2^-15 of the block maximum), packed weight fields are fed to dp4a without being expanded
// ((w & 0x0f0f0f0f) and (w & 0xf0f0f0f0) ARE four byte operands), integer sums are exact, and one fp32 multiply per
// (column, group) applies weight scale x row scale.  Zero points come out as  -zero * sum(row block).
//
// Structure (one CTA per SM, 8 warps, <= 113 KB of shared memory so that TWO consecutive launches are co-resident):
//   * the packed matrix (tcgen05 layout of layout.h: per 32-column block one contiguous byte stream over K) is cut into
//     equal BYTE ranges per CTA, and each CTA's slabs equally over its warps; a warp streams its two adjacent 32-column
//     blocks with cp.async.bulk into a private 3-stage ring and never synchronises with another warp in the main loop.
//   * griddepcontrol.launch_dependents is the first thing a CTA does and every warp requests its first stages BEFORE
//     griddepcontrol.wait: while launch N computes, launch N+1 is already resident and fills its rings, so HBM keeps
//     streaming across the kernel boundary (5.1 TB/s for a decode-shaped chain vs 3.9 TB/s without, same ubench).
//   * after the wait: gather the row through q_perm, apply the prologue op, quantise, stage in shared memory.
//   * split-K: warp partials -> shared memory -> one partial per (CTA, 64-column pair) -> fp32 workspace + arrival
//     counter; the last arriver sums the slots in CTA order (deterministic) and writes fp16 (+bias, +residual).
#include <map>
#include <mutex>

#include "gemv_i8.cuh"

namespace exl2b {

constexpr int I8_WARPS = 8;
constexpr int I8_THREADS = I8_WARPS * 32;
constexpr int I8_MAX_STAGES = 4;
constexpr int I8_PERM_PREFETCH = 2;       // staging rounds whose permutation indices are fetched before the dependency wait

struct I8Mat {
    const uint8_t* packed;
    const uint32_t* q_scale;
    const half* q_scale_max;
    const uint32_t* qzeros;
    const half* gptq_scales;
    const half* bias;
    half* c;
    unsigned long long byte_base;     // bytes of the matrices before this one in the launch
    uint32_t blk_stream_bytes;
    int N, gp_base, is_gptq, clear, num_regions;
    QRegion reg[MAX_REGIONS];
};

struct I8Params {
    I8Mat mat[I8_MAX_MATS];
    int num_mats, K, KS, GP;          // GP: 64-column pairs over all matrices
    const uint16_t* perm;
    const half* x;
    const half* x2;
    const half* norm_w;
    float norm_eps;
    int mode;
    unsigned long long total_bytes;
    float* ws;
    unsigned int* counters;
    int max_slots;
    int slot_bytes, ns;               // weight ring: bytes per stage slot, slots per warp
    int act_cap;                      // staged row capacity in slabs (multiple of 4)
    unsigned char spp[9];             // slabs per stage by bit width
    unsigned long long* dbg;          // optional globaltimer stamps (exl2b_debug_set), NULL in production
    int dbg_cta;
};

// ---- small device helpers ------------------------------------------------------------------------------------------------
__device__ __forceinline__ int dp4a_us(uint32_t a, uint32_t b, int c) {      // a: 4 unsigned bytes, b: 4 signed bytes
    int d;
    asm("dp4a.u32.s32 %0, %1, %2, %3;" : "=r"(d) : "r"(a), "r"(b), "r"(c));
    return d;
}
__device__ __forceinline__ int dp4a_uu(uint32_t a, uint32_t b, int c) {
    int d;
    asm("dp4a.u32.u32 %0, %1, %2, %3;" : "=r"(d) : "r"(a), "r"(b), "r"(c));
    return d;
}

// byte position -> unit containing it
__device__ __forceinline__ int locate(const I8Params& P, unsigned long long pos) {
    int mi = 0;
#pragma unroll
    for (int i = 1; i < I8_MAX_MATS; ++i)
        if (i < P.num_mats && pos >= P.mat[i].byte_base) mi = i;
    const I8Mat& m = P.mat[mi];
    const unsigned long long rel = pos - m.byte_base;
    const uint32_t pair_bytes = 2u * m.blk_stream_bytes;
    const uint32_t p = (uint32_t)(rel / pair_bytes);
    const uint32_t off = (uint32_t)(rel - (unsigned long long)p * pair_bytes) >> 1;
    int r = 0;
#pragma unroll
    for (int i = 1; i < MAX_REGIONS; ++i)
        if (i < m.num_regions && off >= m.reg[i].off_base) r = i;
    const int ks = m.reg[r].ks_begin + (int)((off - m.reg[r].off_base) / (uint32_t)(128 * m.reg[r].bits));
    return (m.gp_base + (int)p) * P.KS + ks;
}

// ---- one slab (32 k) of the warp's two 32-column blocks: integer dot products ---------------------------------------------
// Staged row of a slab (64 B): XH[2j] / XH[2j+1] = high bytes of k = 8j + {0,4,1,5} / 8j + {2,6,3,7}; XL the low bytes.
// That is the byte order (w & 0x0f0f0f0f) / (w & 0xf0f0f0f0) of a 4-bit plane word has (layout.h: field e of pair slot j
// at bit 16e + 4j, k = 8w + 2j + e); other planes reach their order with one PRMT per operand.
// Plain shared-memory loads (not asm volatile): the compiler is free to hoist the next slab's loads above this slab's math.
template <int BITS>
__device__ __forceinline__ void consume_slab(const uint8_t* __restrict__ w0, const uint8_t* __restrict__ w1,
                                             const uint8_t* __restrict__ xs, int lane, int (&am)[2][4], int (&ae)[2][2]) {
    constexpr int Pm = plane_main(BITS), Pe = plane_extra(BITS);
    uint32_t XH[8], XL[8];
    {
        const uint4* xp = reinterpret_cast<const uint4*>(xs);
        const uint4 h0 = xp[0], h1 = xp[1], l0 = xp[2], l1 = xp[3];
        XH[0] = h0.x; XH[1] = h0.y; XH[2] = h0.z; XH[3] = h0.w; XH[4] = h1.x; XH[5] = h1.y; XH[6] = h1.z; XH[7] = h1.w;
        XL[0] = l0.x; XL[1] = l0.y; XL[2] = l0.z; XL[3] = l0.w; XL[4] = l1.x; XL[5] = l1.y; XL[6] = l1.z; XL[7] = l1.w;
    }
    // ---- main plane
    if constexpr (Pm == 4) {
#pragma unroll
        for (int c = 0; c < 2; ++c) {
            const uint4 w4 = reinterpret_cast<const uint4*>(c ? w1 : w0)[lane];
            const uint32_t W[4] = {w4.x, w4.y, w4.z, w4.w};
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                const uint32_t lo = W[j] & 0x0f0f0f0fu, hi = W[j] & 0xf0f0f0f0u;      // hi carries a factor 16 (removed at the flush)
                am[c][0] = dp4a_us(lo, XH[2 * j], am[c][0]);
                am[c][1] = dp4a_uu(lo, XL[2 * j], am[c][1]);
                am[c][2] = dp4a_us(hi, XH[2 * j + 1], am[c][2]);
                am[c][3] = dp4a_uu(hi, XL[2 * j + 1], am[c][3]);
            }
        }
    } else if constexpr (Pm == 8) {
        // word w: bytes = k 4w + {0,2,1,3}
        uint32_t YH[8], YL[8];
#pragma unroll
        for (int w = 0; w < 8; ++w) {
            const int j = w >> 1;
            const uint32_t sel = (w & 1) ? 0x7351u : 0x6240u;
            YH[w] = __byte_perm(XH[2 * j], XH[2 * j + 1], sel);
            YL[w] = __byte_perm(XL[2 * j], XL[2 * j + 1], sel);
        }
#pragma unroll
        for (int c = 0; c < 2; ++c) {
            const uint4* wp = reinterpret_cast<const uint4*>(c ? w1 : w0);
            const uint4 a4 = wp[lane], b4 = wp[32 + lane];
            const uint32_t W[8] = {a4.x, a4.y, a4.z, a4.w, b4.x, b4.y, b4.z, b4.w};
#pragma unroll
            for (int w = 0; w < 8; ++w) {
                am[c][0] = dp4a_us(W[w], YH[w], am[c][0]);
                am[c][1] = dp4a_uu(W[w], YL[w], am[c][1]);
            }
        }
    } else {   // Pm == 2: word w covers k = 16w .. 16w+15; field i of a byte: k 16w + {2i, 8+2i, 2i+1, 9+2i}
        uint32_t YH[8], YL[8];
#pragma unroll
        for (int w = 0; w < 2; ++w)
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                const int a = 2 * w, b = 2 * w + 1, hi = i & 1;
                const uint32_t sel = (i & 2) ? 0x7351u : 0x6240u;
                YH[w * 4 + i] = __byte_perm(XH[2 * a + hi], XH[2 * b + hi], sel);
                YL[w * 4 + i] = __byte_perm(XL[2 * a + hi], XL[2 * b + hi], sel);
            }
#pragma unroll
        for (int c = 0; c < 2; ++c) {
            const uint2 w2 = reinterpret_cast<const uint2*>(c ? w1 : w0)[lane];
            const uint32_t W[2] = {w2.x, w2.y};
#pragma unroll
            for (int w = 0; w < 2; ++w)
#pragma unroll
                for (int i = 0; i < 4; ++i) {
                    const uint32_t t = (W[w] >> (2 * i)) & 0x03030303u;
                    am[c][0] = dp4a_us(t, YH[w * 4 + i], am[c][0]);
                    am[c][1] = dp4a_uu(t, YL[w * 4 + i], am[c][1]);
                }
        }
    }
    // ---- extra plane (bits above the main plane), at byte 128 * Pm of the block
    if constexpr (Pe == 1) {   // one word: bit i of a byte: k = {2i, 16+2i, 2i+1, 17+2i}
        uint32_t YH[8], YL[8];
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            const int oa = i >> 2, ob = 2 + (i >> 2), t = i & 3, hi = t & 1;
            const uint32_t sel = (t & 2) ? 0x7351u : 0x6240u;
            YH[i] = __byte_perm(XH[2 * oa + hi], XH[2 * ob + hi], sel);
            YL[i] = __byte_perm(XL[2 * oa + hi], XL[2 * ob + hi], sel);
        }
#pragma unroll
        for (int c = 0; c < 2; ++c) {
            const uint32_t w = reinterpret_cast<const uint32_t*>((c ? w1 : w0) + 128 * Pm)[lane];
#pragma unroll
            for (int i = 0; i < 8; ++i) {
                const uint32_t t = (w >> i) & 0x01010101u;
                ae[c][0] = dp4a_us(t, YH[i], ae[c][0]);
                ae[c][1] = dp4a_uu(t, YL[i], ae[c][1]);
            }
        }
    } else if constexpr (Pe == 2) {
        uint32_t YH[8], YL[8];
#pragma unroll
        for (int w = 0; w < 2; ++w)
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                const int a = 2 * w, b = 2 * w + 1, hi = i & 1;
                const uint32_t sel = (i & 2) ? 0x7351u : 0x6240u;
                YH[w * 4 + i] = __byte_perm(XH[2 * a + hi], XH[2 * b + hi], sel);
                YL[w * 4 + i] = __byte_perm(XL[2 * a + hi], XL[2 * b + hi], sel);
            }
#pragma unroll
        for (int c = 0; c < 2; ++c) {
            const uint2 w2 = reinterpret_cast<const uint2*>((c ? w1 : w0) + 128 * Pm)[lane];
            const uint32_t W[2] = {w2.x, w2.y};
#pragma unroll
            for (int w = 0; w < 2; ++w)
#pragma unroll
                for (int i = 0; i < 4; ++i) {
                    const uint32_t t = (W[w] >> (2 * i)) & 0x03030303u;
                    ae[c][0] = dp4a_us(t, YH[w * 4 + i], ae[c][0]);
                    ae[c][1] = dp4a_uu(t, YL[w * 4 + i], ae[c][1]);
                }
        }
    }
}

template <int BITS>
__device__ __forceinline__ int consume_stage(const uint8_t* __restrict__ slot, int n, const uint8_t* __restrict__ xs,
                                             const int* __restrict__ asum, int lane, int (&am)[2][4], int (&ae)[2][2]) {
    constexpr int bb = 128 * BITS;
    int S = 0;
    const uint8_t* w1 = slot + n * bb;
#pragma unroll 2
    for (int s = 0; s < n; ++s) {
        consume_slab<BITS>(slot + s * bb, w1 + s * bb, xs + s * 64, lane, am, ae);
        S += asum[s];
    }
    return S;
}

// scale / zero-point words of one (group, 64-column pair), fetched a segment ahead of their use
struct RawScale {
    uint32_t a, b;     // EXL2: q_scale words of the lane's two columns;  GPTQ: qzeros words
    half ha, hb;       // GPTQ: fp16 scales
    half hm;           // EXL2: q_scale_max[group]
};
__device__ __forceinline__ RawScale load_scales(const I8Params& P, int mi, int gp, int group, int lane) {
    const I8Mat& m = P.mat[mi];
    const int nA = (gp - m.gp_base) * 64 + lane, nB = nA + 32;
    const int wpr = m.N >> 3;
    RawScale r;
    r.ha = r.hb = r.hm = __ushort_as_half(0);
    if (!m.is_gptq) {
        r.a = nA < m.N ? __ldg(m.q_scale + (size_t)group * wpr + (nA >> 3)) : 0u;
        r.b = nB < m.N ? __ldg(m.q_scale + (size_t)group * wpr + (nB >> 3)) : 0u;
        r.hm = __ldg(m.q_scale_max + group);
    } else {
        r.a = nA < m.N ? __ldg(m.qzeros + (size_t)group * wpr + (nA >> 3)) : 0u;
        r.b = nB < m.N ? __ldg(m.qzeros + (size_t)group * wpr + (nB >> 3)) : 0u;
        if (nA < m.N) r.ha = __ldg(m.gptq_scales + (size_t)group * m.N + nA);
        if (nB < m.N) r.hb = __ldg(m.gptq_scales + (size_t)group * m.N + nB);
    }
    return r;
}

__device__ __forceinline__ void finalize_pair(const I8Params& P, int gp, int lane, float v0, float v1) {
    int mi = 0;
#pragma unroll
    for (int i = 1; i < I8_MAX_MATS; ++i)
        if (i < P.num_mats && gp >= P.mat[i].gp_base) mi = i;
    const I8Mat& m = P.mat[mi];
    const int nA = (gp - m.gp_base) * 64 + lane, nB = nA + 32;
    if (nA < m.N) {
        float v = v0;
        if (m.bias) v += __half2float(m.bias[nA]);
        if (!m.clear) v += __half2float(m.c[nA]);
        m.c[nA] = __float2half_rn(v);
    }
    if (nB < m.N) {
        float v = v1;
        if (m.bias) v += __half2float(m.bias[nB]);
        if (!m.clear) v += __half2float(m.c[nB]);
        m.c[nB] = __float2half_rn(v);
    }
}

__device__ __forceinline__ half silu_h(half x) {        // cuda/q_mlp_activation.cuh:13-23, same fp16 op sequence
    const half e = hexp(__hneg(x));
    const half r = hrcp(__hadd(__float2half(1.0f), e));
    return __hmul(x, r);
}
__device__ __forceinline__ half gelu_h(half x) {        // cuda/q_mlp_activation.cuh:37-47
    float xf = __half2float(x);
    const float t = 0.797884560803f * (xf + 0.044715f * xf * xf * xf);
    float th;
    asm("tanh.approx.f32 %0, %1;" : "=f"(th) : "f"(t));
    xf = 0.5f * xf * (1.0 + th);
    return __float2half_rn(xf);
}

constexpr int DF_FLUSH = 1, DF_PAIR_DONE = 2;

__device__ __forceinline__ unsigned long long i8_gtimer() {
    unsigned long long t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    return t;
}
// stamps of warp 0 / lane 0 of CTA dbg_cta: 0 start, 1 first stages requested, 2 dependency wait over, 3 row staged,
// 4 warp 0's main loop done, 5 all warps done, (6 = earliest CTA start, 7 = latest CTA end over the grid)
#define I8_STAMP(i) do { if (P.dbg) { if (blockIdx.x == P.dbg_cta && tid == 0) P.dbg[i] = i8_gtimer(); if ((i) == 0 && tid == 0) atomicMin(P.dbg + 6, i8_gtimer()); } } while (0)

// ---- the kernel -----------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(I8_THREADS, 2) gemv_i8_kernel(const __grid_constant__ I8Params P) {
    extern __shared__ __align__(128) uint8_t smem[];
    __shared__ uint64_t bars[I8_WARPS * I8_MAX_STAGES];
    __shared__ int4 descs[I8_WARPS * I8_MAX_STAGES];       // stage descriptors: written at issue, read at consumption
    __shared__ float s_red[I8_WARPS];
    __shared__ int em_gp[I8_WARPS][2], em_n[I8_WARPS][2];

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int KS = P.KS, ns = P.ns;
    I8_STAMP(0);
    if (tid < I8_WARPS * I8_MAX_STAGES) mbar_init(smem_addr(&bars[tid]), 1);
    if (tid < I8_WARPS * 2) em_gp[tid >> 1][tid & 1] = -1;
    mbar_fence_init();
    __syncthreads();
    griddep_launch_dependents();          // the next launch may become resident (and prefetch ITS weights) right away

    // ---- this CTA's / this warp's unit range
    const unsigned C = gridDim.x, cta = blockIdx.x;
    const int L0 = locate(P, P.total_bytes * cta / C);
    const int L1 = (cta + 1 == C) ? P.GP * KS : locate(P, P.total_bytes * (cta + 1) / C);
    const int nunits = L1 - L0;
    if (nunits <= 0) return;
    const int l0 = L0 + (int)((long long)nunits * warp / I8_WARPS), l1 = L0 + (int)((long long)nunits * (warp + 1) / I8_WARPS);

    const uint32_t slot_bytes = (uint32_t)P.slot_bytes;
    uint8_t* const ring_g = smem + (size_t)warp * (size_t)(ns * slot_bytes);
    const uint32_t ring = smem_addr(ring_g);
    const uint32_t bar0 = smem_addr(&bars[warp * I8_MAX_STAGES]);
    int4* const my_descs = descs + warp * I8_MAX_STAGES;
    const uint32_t ring_bytes = (uint32_t)I8_WARPS * (uint32_t)ns * slot_bytes;
    uint8_t* const act_g = smem + ring_bytes;
    int* const asum_s = reinterpret_cast<int*>(act_g + (size_t)P.act_cap * 64);
    float* const ascale_s = reinterpret_cast<float*>(asum_s + P.act_cap);
    float* const emit_base = reinterpret_cast<float*>(reinterpret_cast<uint8_t*>(ascale_s) + (size_t)(P.act_cap / 4 + 1) * 4);   // [warp][2][64]

    // ---- which slabs of the row this CTA needs (whole 128-k blocks), as at most two intervals A (first pair) and B (last pair)
    const int gpA = L0 / KS, gpB = (L1 - 1) / KS;
    int a0, a1, b1;
    {
        const int a_ks = L0 - gpA * KS, b_end = (L1 - 1) - gpB * KS + 1;
        if (gpA == gpB) {
            a0 = a_ks & ~3; a1 = min(KS, (b_end + 3) & ~3); b1 = 0;
        } else if (gpB == gpA + 1) {
            a0 = a_ks & ~3; a1 = KS; b1 = min(KS, (b_end + 3) & ~3);
            if (b1 >= a0) { a0 = 0; b1 = 0; }
        } else {
            a0 = 0; a1 = KS; b1 = 0;
        }
    }
    const int lenA = a1 - a0, lenA_pad = (lenA + 3) & ~3, lenB = b1;
    const bool whole = (a0 == 0 && a1 == KS);
    const int n_oct = (lenA_pad + ((lenB + 3) & ~3)) * 4;
    if (lenA_pad + ((lenB + 3) & ~3) > P.act_cap) __trap();

    // ---- issue cursor: position (i_gp, i_ks), matrix i_mi and its current region cached in registers
    int i_lin = l0, i_gp = l0 / KS, i_ks = l0 - (l0 / KS) * KS, i_mi = 0, i_r = 0, islot = 0;
#pragma unroll
    for (int i = 1; i < I8_MAX_MATS; ++i)
        if (i < P.num_mats && i_gp >= P.mat[i].gp_base) i_mi = i;
    int r_begin, r_bits, r_spg, r_gbase, r_end;
    uint32_t r_off;
    const uint8_t* i_src;        // block stream of the even block of pair i_gp
    uint32_t i_bsb;              // bytes of one block stream of matrix i_mi
    auto load_region = [&]() {
        const I8Mat& m = P.mat[i_mi];
        const QRegion& rg = m.reg[i_r];
        r_begin = rg.ks_begin; r_bits = rg.bits; r_spg = rg.spg_log2; r_gbase = rg.group_base; r_off = rg.off_base;
        r_end = (i_r + 1 < m.num_regions) ? m.reg[i_r + 1].ks_begin : KS;
    };
    auto load_pair = [&]() {
        const I8Mat& m = P.mat[i_mi];
        i_bsb = m.blk_stream_bytes;
        i_src = m.packed + (size_t)(2 * (i_gp - m.gp_base)) * i_bsb;
    };
    {
        const I8Mat& m = P.mat[i_mi];
#pragma unroll
        for (int i = 1; i < MAX_REGIONS; ++i)
            if (i < m.num_regions && i_ks >= m.reg[i].ks_begin) i_r = i;
    }
    load_region();
    load_pair();
    auto issue_one = [&]() {
        const int rel = i_ks - r_begin, g = rel >> r_spg;
        const int gend = r_begin + ((g + 1) << r_spg);
        const int segend = min(min(gend, (i_ks | 3) + 1), min(r_end, i_ks + (l1 - i_lin)));
        const int n = min(segend - i_ks, r_bits > 4 ? 2 : 4);          // <= 4 KB per stage (two copies of n * 128 * bits bytes)
        if (lane == 0) {
            const uint32_t bar = bar0 + islot * 8;
            const uint32_t bytes = (uint32_t)(n * 128 * r_bits);
            const uint8_t* src = i_src + r_off + (uint32_t)(rel * 128 * r_bits);
            const uint32_t dst = ring + (uint32_t)islot * slot_bytes;
            mbar_arrive_expect_tx(bar, 2 * bytes);
            bulk_copy_g2s(dst, src, bytes, bar);
            bulk_copy_g2s(dst + bytes, src + i_bsb, bytes, bar);
            const int flags = ((i_ks + n == segend) ? DF_FLUSH : 0) | ((i_ks + n == KS || i_lin + n == l1) ? DF_PAIR_DONE : 0);
            const int si0 = (i_gp == gpA || whole) ? i_ks - a0 : lenA_pad + i_ks;
            my_descs[islot] = make_int4(n | (r_bits << 8) | (flags << 16), si0, r_gbase + g, i_gp | (i_mi << 24));
        }
        i_lin += n;
        i_ks += n;
        islot = (islot + 1 == ns) ? 0 : islot + 1;
        if (i_ks >= KS) {
            i_ks = 0;
            i_gp++;
            if (i_mi + 1 < P.num_mats && i_gp >= P.mat[i_mi + 1].gp_base) i_mi++;
            i_r = 0;
            load_region();
            load_pair();
        } else if (i_ks >= r_end) {
            i_r++;
            load_region();
        }
    };
    // weight requests for the first stages (nothing here depends on the previous launch)
    for (int i = 0; i < ns && i_lin < l1; ++i) issue_one();
    __syncwarp();
    RawScale raw = {};
    if (l0 < l1) {
        const int4 d0 = my_descs[0];
        raw = load_scales(P, d0.w >> 24, d0.w & 0xffffff, d0.z, lane);
    }

    auto oct_k0 = [&](int i, bool& valid) -> int {      // staged octet i -> first k of its 8 (and whether it exists)
        const int si = i >> 2, j = i & 3;
        int ks;
        if (si < lenA_pad) { ks = a0 + si; valid = (i < n_oct) && si < lenA; }
        else { ks = si - lenA_pad; valid = (i < n_oct) && (si - lenA_pad) < lenB; }
        return ks * 32 + j * 8;
    };
    uint4 pv[I8_PERM_PREFETCH];
#pragma unroll
    for (int r = 0; r < I8_PERM_PREFETCH; ++r) {
        bool valid;
        const int k0 = oct_k0(r * I8_THREADS + tid, valid);
        pv[r] = make_uint4(0, 0, 0, 0);
        if (valid && P.perm) pv[r] = __ldg(reinterpret_cast<const uint4*>(P.perm + k0));
    }

    I8_STAMP(1);
    griddep_wait();                        // everything below may read what the previous launch wrote
    I8_STAMP(2);

    // ---- prologue: the row -> (optional RMSNorm / act*mul) -> 16-bit integers per 128-k block -> shared memory
    float rrms = 1.f;
    if (P.mode == I8_RMSNORM) {
        float sum = 0.f;
        for (int k = tid * 8; k < P.K; k += I8_THREADS * 8) {
            const uint4 v = __ldcg(reinterpret_cast<const uint4*>(P.x + k));
            const half2* h = reinterpret_cast<const half2*>(&v);
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                const float f0 = fmaxf(-65504.f, fminf(__low2float(h[i]), 65504.f));
                const float f1 = fmaxf(-65504.f, fminf(__high2float(h[i]), 65504.f));
                sum = fmaf(f0, f0, sum);
                sum = fmaf(f1, f1, sum);
            }
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, o);
        if (lane == 0) s_red[warp] = sum;
        __syncthreads();
        sum = 0.f;
#pragma unroll
        for (int i = 0; i < I8_WARPS; ++i) sum += s_red[i];
        rrms = rsqrtf(sum * (1.0f / (float)P.K) + P.norm_eps);
    }
    auto stage_round = [&](int i, uint4 pidx) {
        bool valid;
        const int k0 = oct_k0(i, valid);
        float f[8];
#pragma unroll
        for (int e = 0; e < 8; ++e) f[e] = 0.f;
        if (valid) {
            const uint16_t* pi = reinterpret_cast<const uint16_t*>(&pidx);
            int src[8];
            half h[8];
#pragma unroll
            for (int e = 0; e < 8; ++e) src[e] = P.perm ? (int)pi[e] : k0 + e;
#pragma unroll
            for (int e = 0; e < 8; ++e) h[e] = __ldcg(P.x + src[e]);
            if (P.mode == I8_RMSNORM) {
#pragma unroll
                for (int e = 0; e < 8; ++e) {
                    const float xf = fmaxf(-65504.f, fminf(__half2float(h[e]), 65504.f));
                    h[e] = __float2half_rn(xf * __half2float(__ldg(P.norm_w + src[e])) * rrms);
                }
            } else if (P.mode == I8_SILU_MUL) {
#pragma unroll
                for (int e = 0; e < 8; ++e) h[e] = __hmul(silu_h(h[e]), __ldcg(P.x2 + src[e]));
            } else if (P.mode == I8_GELU_MUL) {
#pragma unroll
                for (int e = 0; e < 8; ++e) h[e] = __hmul(gelu_h(h[e]), __ldcg(P.x2 + src[e]));
            }
#pragma unroll
            for (int e = 0; e < 8; ++e) f[e] = __half2float(h[e]);
        }
        float amax = 0.f;
#pragma unroll
        for (int e = 0; e < 8; ++e) amax = fmaxf(amax, fabsf(f[e]));
#pragma unroll
        for (int o = 1; o < 16; o <<= 1) amax = fmaxf(amax, __shfl_xor_sync(0xffffffffu, amax, o));      // 16 lanes = one 128-k block
        // power-of-two block scale: |q| <= 2^15 - 16, every row value within a factor 16 of the block maximum is represented
        // EXACTLY (fp16 has 11 significant bits), and scale products stay exact -- a unit-vector row returns reconstruct()'s
        // fp16 weights bit for bit
        const uint32_t ef = (__float_as_uint(amax) >> 23) & 0xffu;
        const float inv = amax > 0.f ? __uint_as_float((268u - ef) << 23) : 0.f;
        int q[8], sum = 0;
#pragma unroll
        for (int e = 0; e < 8; ++e) { q[e] = __float2int_rn(f[e] * inv); sum += q[e]; }
        sum += __shfl_xor_sync(0xffffffffu, sum, 1);
        sum += __shfl_xor_sync(0xffffffffu, sum, 2);                                                       // 4 lanes = one slab
        if (valid) {
            auto pack = [&](int i0, int i1, int i2, int i3, int sh) -> uint32_t {
                return ((uint32_t)(q[i0] >> sh) & 0xffu) | (((uint32_t)(q[i1] >> sh) & 0xffu) << 8) |
                       (((uint32_t)(q[i2] >> sh) & 0xffu) << 16) | (((uint32_t)(q[i3] >> sh) & 0xffu) << 24);
            };
            const int si = i >> 2, j = i & 3;
            uint2 hw, lw;
            hw.x = pack(0, 4, 1, 5, 8); hw.y = pack(2, 6, 3, 7, 8);
            lw.x = pack(0, 4, 1, 5, 0); lw.y = pack(2, 6, 3, 7, 0);
            *reinterpret_cast<uint2*>(act_g + (size_t)si * 64 + j * 8) = hw;
            *reinterpret_cast<uint2*>(act_g + (size_t)si * 64 + 32 + j * 8) = lw;
            if (j == 0) asum_s[si] = sum;
            if ((i & 15) == 0) ascale_s[si >> 2] = amax > 0.f ? __uint_as_float((ef - 14u) << 23) : 0.f;
        }
    };
#pragma unroll
    for (int r = 0; r < I8_PERM_PREFETCH; ++r)
        if (r * I8_THREADS < n_oct) stage_round(r * I8_THREADS + tid, pv[r]);
    for (int ob = I8_PERM_PREFETCH * I8_THREADS; ob < n_oct; ob += I8_THREADS) {
        bool valid;
        const int k0 = oct_k0(ob + tid, valid);
        uint4 pidx = make_uint4(0, 0, 0, 0);
        if (valid && P.perm) pidx = __ldg(reinterpret_cast<const uint4*>(P.perm + k0));
        stage_round(ob + tid, pidx);
    }
    __syncthreads();
    I8_STAMP(3);

    // ---- main loop: this warp alone, stage by stage.  Everything positional comes from the descriptor written at issue time.
    int am[2][4] = {{0, 0, 0, 0}, {0, 0, 0, 0}}, ae[2][2] = {{0, 0}, {0, 0}};
    float tot[2] = {0.f, 0.f};
    int S = 0, cslot = 0, pair_slabs = 0, emits = 0, c_lin = l0;
    uint32_t phase = 0;
    while (c_lin < l1) {
        mbar_wait(bar0 + cslot * 8, (phase >> cslot) & 1u);
        phase ^= 1u << cslot;
        const int4 d = my_descs[cslot];
        const int n = d.x & 0xff, bits = (d.x >> 8) & 0xff, flags = d.x >> 16, si0 = d.y;
        const uint8_t* slot = ring_g + (size_t)cslot * slot_bytes;
        const uint8_t* xs = act_g + (size_t)si0 * 64;
        switch (bits) {
            case 4: S += consume_stage<4>(slot, n, xs, asum_s + si0, lane, am, ae); break;
            case 5: S += consume_stage<5>(slot, n, xs, asum_s + si0, lane, am, ae); break;
            case 6: S += consume_stage<6>(slot, n, xs, asum_s + si0, lane, am, ae); break;
            case 3: S += consume_stage<3>(slot, n, xs, asum_s + si0, lane, am, ae); break;
            case 8: S += consume_stage<8>(slot, n, xs, asum_s + si0, lane, am, ae); break;
            default: S += consume_stage<2>(slot, n, xs, asum_s + si0, lane, am, ae); break;
        }
        __syncwarp();
        if (i_lin < l1) issue_one();            // refill the slot just drained (islot == cslot here)
        __syncwarp();
        c_lin += n;
        pair_slabs += n;
        cslot = (cslot + 1 == ns) ? 0 : cslot + 1;
        if (flags & DF_FLUSH) {
            // integer sums -> fp32:  sum_k a_k (q_k - zero) * scale  =  (sum a q - zero * sum a) * scale_w * scale_row
            const int mi = d.w >> 24, gp = d.w & 0xffffff;
            const bool gptq = P.mat[mi].is_gptq != 0;
            const float as = ascale_s[si0 >> 2];
            const int pm = plane_main(bits);
            const int nA = (gp - P.mat[mi].gp_base) * 64 + lane;
#pragma unroll
            for (int c = 0; c < 2; ++c) {
                int v = ((am[c][0] << 8) + am[c][1]) + (((am[c][2] << 8) + am[c][3]) >> 4) + (((ae[c][0] << 8) + ae[c][1]) << pm);
                const uint32_t word = c ? raw.b : raw.a;
                const uint32_t nib = (word >> (((nA + 32 * c) & 7) * 4)) & 15u;
                float ws;
                int zero;
                if (!gptq) {
                    const int qs = (int)nib + 1;
                    ws = __half2float(__hmul(__int2half_rn(qs * qs), raw.hm));          // dq_scale, cuda/quant/qdq_util.cuh:24-30
                    zero = 1 << (bits - 1);
                } else {
                    ws = __half2float(c ? raw.hb : raw.ha);
                    zero = (int)nib + 1;                                                 // q_gemm_kernel_gptq.cuh:167-172
                }
                v -= zero * S;
                tot[c] = fmaf((float)v, ws * as, tot[c]);
                am[c][0] = am[c][1] = am[c][2] = am[c][3] = 0;
                ae[c][0] = ae[c][1] = 0;
            }
            S = 0;
            if (c_lin < l1) {                  // scales of the next segment (its descriptor is already in the ring)
                const int4 dn = my_descs[cslot];
                raw = load_scales(P, dn.w >> 24, dn.w & 0xffffff, dn.z, lane);
            }
            if (flags & DF_PAIR_DONE) {
                if (pair_slabs == KS) {
                    finalize_pair(P, gp, lane, tot[0], tot[1]);          // this warp covered the pair's whole K by itself
                } else {
                    if (emits >= 2) __trap();       // a warp's range has at most two partial pairs (its first and its last)
                    float* dst = emit_base + (warp * 2 + emits) * 64;
                    dst[lane] = tot[0];
                    dst[32 + lane] = tot[1];
                    if (lane == 0) { em_gp[warp][emits] = gp; em_n[warp][emits] = pair_slabs; }
                    ++emits;
                }
                tot[0] = tot[1] = 0.f;
                pair_slabs = 0;
            }
        }
    }
    I8_STAMP(4);
    __syncthreads();
    I8_STAMP(5);

    // ---- CTA-level reduction of the partial pairs, then workspace hand-off (deterministic order everywhere)
    for (int gp = gpA + warp; gp <= gpB; gp += I8_WARPS) {
        float v0 = 0.f, v1 = 0.f;
        int cnt = 0;
#pragma unroll
        for (int w = 0; w < I8_WARPS; ++w)
#pragma unroll
            for (int e = 0; e < 2; ++e)
                if (em_gp[w][e] == gp) {
                    v0 += emit_base[(w * 2 + e) * 64 + lane];
                    v1 += emit_base[(w * 2 + e) * 64 + 32 + lane];
                    cnt += em_n[w][e];
                }
        if (cnt == 0) continue;
        if (cnt == KS) { finalize_pair(P, gp, lane, v0, v1); continue; }
        int mi = 0;
#pragma unroll
        for (int i = 1; i < I8_MAX_MATS; ++i)
            if (i < P.num_mats && gp >= P.mat[i].gp_base) mi = i;
        const I8Mat& m = P.mat[mi];
        const unsigned long long pair_bytes = 2ull * m.blk_stream_bytes;
        const unsigned long long b0 = m.byte_base + (unsigned long long)(gp - m.gp_base) * pair_bytes;
        const unsigned c_first = (unsigned)(((b0 + 256ull * m.reg[0].bits) * C - 1) / P.total_bytes);
        const unsigned c_last = min(C - 1, (unsigned)(((b0 + pair_bytes) * C - 1) / P.total_bytes));
        float* wsp = P.ws + ((size_t)gp * P.max_slots) * 64;
        const unsigned my = cta - c_first;
        __stcg(wsp + (size_t)my * 64 + lane, v0);
        __stcg(wsp + (size_t)my * 64 + 32 + lane, v1);
        __threadfence();
        __syncwarp();
        unsigned old = 0;
        if (lane == 0) {
            old = atomicAdd(P.counters + gp, (unsigned)cnt);
            __threadfence();
        }
        old = __shfl_sync(0xffffffffu, old, 0);
        if (old + (unsigned)cnt == (unsigned)KS) {
            v0 = v1 = 0.f;
            for (unsigned s = 0; s <= c_last - c_first; ++s) {
                v0 += __ldcg(wsp + (size_t)s * 64 + lane);
                v1 += __ldcg(wsp + (size_t)s * 64 + 32 + lane);
            }
            finalize_pair(P, gp, lane, v0, v1);
            if (lane == 0) P.counters[gp] = 0;
        }
    }
    if (P.dbg && lane == 0) atomicMax(P.dbg + 7, i8_gtimer());
}

// ---- host side ---------------------------------------------------------------------------------------------------------------

struct I8Ctx {
    float* ws = nullptr;
    unsigned int* counters = nullptr;
    size_t ws_floats = 0;
    int n_counters = 0;
};
static std::mutex g_i8_mutex;
static std::map<std::pair<int, cudaStream_t>, I8Ctx> g_i8_ctx;      // split-K workspace per (device, stream): launches on
static bool g_i8_attr[64] = {false};                                  // different streams never share scratch

static int i8_ctx(int device, cudaStream_t stream, I8Ctx** out) {
    std::lock_guard<std::mutex> lk(g_i8_mutex);
    I8Ctx& c = g_i8_ctx[{device, stream}];
    if (!c.ws) {
        c.ws_floats = (size_t)4 << 20;        // 16 MB
        c.n_counters = 1 << 15;
        EXL2B_CUDA(cudaMalloc(&c.ws, c.ws_floats * sizeof(float)));
        EXL2B_CUDA(cudaMalloc(&c.counters, c.n_counters * sizeof(unsigned)));
        EXL2B_CUDA(cudaMemsetAsync(c.counters, 0, c.n_counters * sizeof(unsigned), stream));
    }
    if (device >= 0 && device < 64 && !g_i8_attr[device]) {
        EXL2B_CUDA(cudaFuncSetAttribute(gemv_i8_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
        g_i8_attr[device] = true;
    }
    *out = &c;
    return 0;
}

bool gemv_i8_enabled() {
    static const bool on = [] {
        const char* e = getenv("EXL2B_GEMV");
        return !(e && e[0] == 't');
    }();
    return on;
}

bool gemv_i8_fusable(const QMatrix* const* qs, int nm) {
    if (nm < 1 || nm > I8_MAX_MATS) return false;
    for (int i = 0; i < nm; ++i) {
        if (!qs[i] || qs[i]->v.layout != LAYOUT_TC || qs[i]->v.K != qs[0]->v.K || qs[i]->device != qs[0]->device) return false;
        if ((qs[i]->v.perm == nullptr) != (qs[0]->v.perm == nullptr)) return false;
    }
    if (nm == 1 || !qs[0]->v.perm) return true;
    const int K = qs[0]->v.K;
    std::vector<uint16_t> p0(K), pi(K);
    if (cudaSetDevice(qs[0]->device) != cudaSuccess) return false;
    if (cudaMemcpy(p0.data(), qs[0]->v.perm, (size_t)K * 2, cudaMemcpyDeviceToHost) != cudaSuccess) return false;
    for (int i = 1; i < nm; ++i) {
        if (qs[i]->v.perm == qs[0]->v.perm) continue;
        if (cudaMemcpy(pi.data(), qs[i]->v.perm, (size_t)K * 2, cudaMemcpyDeviceToHost) != cudaSuccess) return false;
        if (p0 != pi) return false;
    }
    return true;
}

int gemv_i8_launch(int device, cudaStream_t stream, const I8Out* outs, int nm, const I8Input& in) {
    EXL2B_REQUIRE(nm >= 1 && nm <= I8_MAX_MATS, "bad matrix count %d", nm);
    EXL2B_REQUIRE(in.x, "null input row");
    EXL2B_REQUIRE(in.mode != I8_RMSNORM || in.norm_w, "RMSNorm prologue without a weight");
    EXL2B_REQUIRE((in.mode != I8_SILU_MUL && in.mode != I8_GELU_MUL) || in.x2, "act*mul prologue without the second operand");
    I8Ctx* ctx = nullptr;
    int rc = i8_ctx(device, stream, &ctx);
    if (rc) return rc;

    I8Params P = {};
    P.num_mats = nm;
    P.K = outs[0].q->v.K;
    P.KS = outs[0].q->v.KS;
    P.perm = outs[0].q->v.perm;
    P.x = in.x;
    P.x2 = in.x2;
    P.norm_w = in.norm_w;
    P.norm_eps = in.norm_eps;
    P.mode = in.mode;
    unsigned long long bytes = 0;
    int gp = 0, max_bits = 2, min_bits = 8;
    uint32_t max_stream = 0;
    for (int i = 0; i < nm; ++i) {
        const QMatrix* q = outs[i].q;
        EXL2B_REQUIRE(q && outs[i].c, "null matrix / output");
        const QMatView& v = q->v;
        EXL2B_REQUIRE(v.layout == LAYOUT_TC, "matrix is not in the tcgen05 layout");
        EXL2B_REQUIRE(v.KS == P.KS, "fused matrices must share K");
        EXL2B_REQUIRE((v.perm == nullptr) == (P.perm == nullptr), "fused matrices must share their row permutation");
        I8Mat& m = P.mat[i];
        m.packed = reinterpret_cast<const uint8_t*>(v.packed);
        m.q_scale = v.q_scale;
        m.q_scale_max = v.q_scale_max;
        m.qzeros = v.qzeros;
        m.gptq_scales = v.gptq_scales;
        m.bias = v.bias;
        m.c = outs[i].c;
        m.clear = outs[i].clear;
        m.byte_base = bytes;
        m.blk_stream_bytes = v.blk_stream_bytes;
        m.N = v.N;
        m.gp_base = gp;
        m.is_gptq = v.is_gptq;
        m.num_regions = v.num_regions;
        for (int r = 0; r < v.num_regions; ++r) {
            m.reg[r] = v.reg[r];
            max_bits = std::max(max_bits, v.reg[r].bits);
            min_bits = std::min(min_bits, v.reg[r].bits);
        }
        max_stream = std::max(max_stream, v.blk_stream_bytes);
        gp += v.strips * 2;
        bytes += (unsigned long long)v.strips * 4ull * v.blk_stream_bytes;
    }
    P.GP = gp;
    P.total_bytes = bytes;
    EXL2B_REQUIRE((long long)P.GP * P.KS < (1ll << 30), "problem too large for 32-bit unit arithmetic");

    const int sms = device_sm_count(device);
    const unsigned long long max_unit = 256ull * max_bits, min_unit = 256ull * min_bits;
    const int C = (int)std::max<unsigned long long>(1, std::min<unsigned long long>((unsigned long long)sms, bytes / max_unit));
    P.slot_bytes = 4096;
    for (int b = 0; b <= 8; ++b) P.spp[b] = 1;
    for (int b : {2, 3, 4, 5, 6, 8}) P.spp[b] = (unsigned char)std::max(1, std::min(4, P.slot_bytes / (256 * b)));
    const int KS_pad = (P.KS + 3) & ~3;
    const long long per_cta_units = (long long)((bytes / C + max_unit + min_unit - 1) / min_unit) + 2;
    // up to two intervals, each widened to 128-k blocks and padded
    P.act_cap = (int)std::min<long long>(KS_pad, ((per_cta_units + 3) & ~3ll) + 12);
    const size_t act_bytes = (size_t)P.act_cap * 64 + (size_t)P.act_cap * 4 + (size_t)(P.act_cap / 4 + 1) * 4;
    const size_t emit_bytes = (size_t)I8_WARPS * 2 * 64 * sizeof(float);
    P.ns = 3;
    auto smem_for = [&](int ns) { return (size_t)I8_WARPS * ns * P.slot_bytes + act_bytes + emit_bytes; };
    // two launches co-resident per SM (227 KB, 1 KB reserved per CTA) is what lets the next launch prefetch: keep <= 112 KB if a
    // 2-stage ring achieves it, never go below 2 stages
    if (smem_for(P.ns) > 112 * 1024 && smem_for(2) <= 112 * 1024) P.ns = 2;
    const size_t smem_total = smem_for(P.ns);
    EXL2B_REQUIRE(smem_total <= 200 * 1024, "shared memory budget exceeded (%zu bytes, K = %d)", smem_total, P.K);
    P.max_slots = (int)((2ull * max_stream * (unsigned long long)C) / bytes) + 2;
    EXL2B_REQUIRE(P.GP <= ctx->n_counters, "too many column pairs (%d) for the counter array", P.GP);
    EXL2B_REQUIRE((size_t)P.GP * P.max_slots * 64 <= ctx->ws_floats, "split-K workspace too small");
    P.ws = ctx->ws;
    P.counters = ctx->counters;
    extern unsigned long long* g_dbg;
    extern int g_dbg_cta, g_dbg_slot;
    P.dbg = g_dbg ? g_dbg + 32 * (g_dbg_slot++ % 64) : nullptr;
    P.dbg_cta = g_dbg_cta;
    EXL2B_CUDA(launch_pdl(gemv_i8_kernel, dim3(C), dim3(I8_THREADS), smem_total, stream, P));
    return 0;
}

}  // namespace exl2b
