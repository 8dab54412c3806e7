// gemv_i8_kernel -- the batch-1 (decode) dequant-GEMV of libexl2b200, HBM-bound by design.
// Replaces, for one input row, gemm_half_q_half_kernel / gemm_half_q_half_gptq_kernel (exllamav2_ext/cuda/q_gemm_kernel.cuh:
// 140-565, q_gemm_kernel_gptq.cuh:60-225) together with the kernels the reference runs in front of them (rms_norm_kernel,
// cuda/rms_norm.cu:55-143; act_mul_kernel, cuda/q_mlp_activation.cuh:54-100), which here are the kernel's PROLOGUE.
//
// Why integers: measured on B200 (tools/ubench/dp4a.cu, profiles/r02_ubench_dp4a_pdlchain.txt) the reference's recipe
// (unpack to fp16, HFMA2) tops out at 38-44 4-bit weights/clk/SM while HBM delivers 45; the same loop on the integer
// dot-product instruction (IDP.4A) runs at 75-98.  So the row is quantised ONCE per launch to 16-bit integers per 128-k
// block with a power-of-two scale (signed high byte plane + unsigned low byte plane; values within 16x of the block
// maximum are exact, the rest carry |error| <= 2^-15 of the block maximum), packed weight fields are fed to dp4a without
// being expanded ((w & 0x0f0f0f0f) and (w & 0xf0f0f0f0) ARE four byte operands), integer sums are exact, and one fp32
// multiply per (column, group) applies weight scale x row scale.  Zero points come out as  -zero * sum(row block).
//
// Why this shape: a decode step is a chain of ~160 DEPENDENT launches of 8-45 MB each; what bounds it is the serial latency
// of every launch (measured with the phase stamps below), not instruction throughput.  So:
//   * one CTA per SM, 16 warps, <= 112 KB of shared memory: TWO consecutive launches are co-resident.  A CTA's first action is
//     griddepcontrol.launch_dependents and every warp requests its first weight stages BEFORE griddepcontrol.wait, so while
//     launch N computes, launch N+1 is already filling its rings and HBM keeps streaming across the kernel boundary
//     (5.1 TB/s for a decode-shaped chain of dependent launches vs 3.9 TB/s in plain stream order, tools/ubench/pdlchain.cu).
//   * a CTA owns WHOLE 32-column blocks (all of K), its 16 warps split the blocks' K range between them: split-K never leaves
//     the CTA (shared memory + one barrier), there is no workspace, no atomics, no fence, and summation order is fixed.
//     The block -> CTA table is computed on the host and travels in the kernel parameters (no divisions on the device).
//   * a warp streams its block's bytes (tcgen05 layout of layout.h: one contiguous stream per block) with cp.async.bulk into a
//     private 3-stage ring and never synchronises with another warp in the main loop.
//   * when the producer of the row scattered a copy in this matrix's stored-row order (I8Out::c_perm), the prologue reads the
//     row with one 16-byte load per thread instead of eight 2-byte gathers.
#include <algorithm>
#include <vector>

#include "gemv_i8.cuh"

namespace exl2b {

constexpr int I8_MAX_WARPS = 16;
constexpr int I8_MAX_STAGES = 4;
constexpr int I8_MAX_CTAS = 160;
constexpr int I8_SLOT_BYTES = 2048;       // one stage: 4 slabs (128 k) at <= 4 bits, 2 slabs above

struct I8Mat {
    const uint8_t* packed;
    const uint32_t* q_scale;
    const half* q_scale_max;
    const uint32_t* qzeros;
    const half* gptq_scales;
    const half* bias;
    half* c;
    half* c_perm;
    const uint16_t* out_invperm;
    uint32_t blk_stream_bytes;
    int N, blk_base, is_gptq, clear, num_regions;
    QRegion reg[MAX_REGIONS];
};

struct I8Params {
    I8Mat mat[I8_MAX_MATS];
    int num_mats, K, KS;
    const uint16_t* perm;             // stored row k' <- feature perm[k'], or NULL
    const half* x;
    const half* x2;
    const half* norm_w;
    float norm_eps;
    int mode, x_permuted;
    int ns;                           // weight ring slots per warp
    int lcap;                         // capacity of a warp's stage list
    int nb_max;                       // most blocks any CTA owns
    unsigned long long* dbg;          // optional globaltimer stamps (exl2b_debug_set), NULL in production
    int dbg_cta;
    unsigned short cta_blk[I8_MAX_CTAS + 1];      // CTA c owns 32-column blocks [cta_blk[c], cta_blk[c+1]) of the launch
};

// dynamic shared-memory map of a CTA (byte offsets, every region 16-byte aligned) -- one definition for host and device
struct I8Smem {
    uint32_t act, asum, ascale, emit, list, blksrc, total;
};
__host__ __device__ inline I8Smem i8_smem_map(int warps, int ns, int KS, int lcap, int nb_max) {
    auto up = [](uint32_t x) { return (x + 15u) & ~15u; };
    I8Smem m;
    m.act = up((uint32_t)warps * (uint32_t)ns * I8_SLOT_BYTES);      // staged row: [KS][64 B]
    m.asum = up(m.act + (uint32_t)KS * 64u);                          // [KS] integer sum of a slab's row values
    m.ascale = up(m.asum + (uint32_t)KS * 4u);                        // [KS/4 + 1] scale of a 128-k block
    m.emit = up(m.ascale + (uint32_t)(KS / 4 + 1) * 4u);              // [warp][2][32] partial sums of split blocks
    m.list = up(m.emit + (uint32_t)warps * 256u);                     // [warp][lcap] stage descriptors
    m.blksrc = up(m.list + (uint32_t)warps * (uint32_t)lcap * 8u);    // [nb_max] block streams
    m.total = up(m.blksrc + (uint32_t)nb_max * 8u);
    return m;
}

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

// ---- one slab (32 k) of the warp's 32-column block: integer dot products, one column per lane -------------------------------
// Staged row of a slab (64 B): XH[2j] / XH[2j+1] = high bytes of k = 8j + {0,4,1,5} / 8j + {2,6,3,7}; XL the low bytes.
// That is the byte order (w & 0x0f0f0f0f) / (w & 0xf0f0f0f0) of a 4-bit plane word has (layout.h: field e of pair slot j
// at bit 16e + 4j, k = 8w + 2j + e); other planes reach their order with one PRMT per operand.
template <int BITS>
__device__ __forceinline__ void consume_slab(uint32_t wb, uint32_t xs, int lane, int (&am)[4], int (&ae)[2]) {
    constexpr int Pm = plane_main(BITS), Pe = plane_extra(BITS);
    if constexpr (BITS == 4) {
        // the common case, in two halves of the slab so that only 8 row words are live at a time (64 registers per thread)
        const uint4 w4 = lds128(wb + lane * 16);
#pragma unroll
        for (int hf = 0; hf < 2; ++hf) {
            const uint4 xh = lds128(xs + hf * 16), xl = lds128(xs + 32 + hf * 16);
            const uint32_t wa = hf ? w4.z : w4.x, wc = hf ? w4.w : w4.y;
            const uint32_t lo0 = wa & 0x0f0f0f0fu, hi0 = wa & 0xf0f0f0f0u, lo1 = wc & 0x0f0f0f0fu, hi1 = wc & 0xf0f0f0f0u;
            am[0] = dp4a_us(lo0, xh.x, am[0]);
            am[1] = dp4a_uu(lo0, xl.x, am[1]);
            am[2] = dp4a_us(hi0, xh.y, am[2]);
            am[3] = dp4a_uu(hi0, xl.y, am[3]);
            am[0] = dp4a_us(lo1, xh.z, am[0]);
            am[1] = dp4a_uu(lo1, xl.z, am[1]);
            am[2] = dp4a_us(hi1, xh.w, am[2]);
            am[3] = dp4a_uu(hi1, xl.w, am[3]);
        }
        return;
    }
    uint32_t XH[8], XL[8];
    {
        const uint4 h0 = lds128(xs), h1 = lds128(xs + 16), l0 = lds128(xs + 32), l1 = lds128(xs + 48);
        XH[0] = h0.x; XH[1] = h0.y; XH[2] = h0.z; XH[3] = h0.w; XH[4] = h1.x; XH[5] = h1.y; XH[6] = h1.z; XH[7] = h1.w;
        XL[0] = l0.x; XL[1] = l0.y; XL[2] = l0.z; XL[3] = l0.w; XL[4] = l1.x; XL[5] = l1.y; XL[6] = l1.z; XL[7] = l1.w;
    }
    // ---- main plane
    if constexpr (Pm == 4) {
        const uint4 w4 = lds128(wb + lane * 16);
        const uint32_t W[4] = {w4.x, w4.y, w4.z, w4.w};
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const uint32_t lo = W[j] & 0x0f0f0f0fu, hi = W[j] & 0xf0f0f0f0u;      // hi carries a factor 16 (removed at the flush)
            am[0] = dp4a_us(lo, XH[2 * j], am[0]);
            am[1] = dp4a_uu(lo, XL[2 * j], am[1]);
            am[2] = dp4a_us(hi, XH[2 * j + 1], am[2]);
            am[3] = dp4a_uu(hi, XL[2 * j + 1], am[3]);
        }
    } else if constexpr (Pm == 8) {
        // word w: bytes = k 4w + {0,2,1,3}
        const uint4 a4 = lds128(wb + lane * 16), b4 = lds128(wb + 512 + lane * 16);
        const uint32_t W[8] = {a4.x, a4.y, a4.z, a4.w, b4.x, b4.y, b4.z, b4.w};
#pragma unroll
        for (int w = 0; w < 8; ++w) {
            const int j = w >> 1;
            const uint32_t sel = (w & 1) ? 0x7351u : 0x6240u;
            am[0] = dp4a_us(W[w], __byte_perm(XH[2 * j], XH[2 * j + 1], sel), am[0]);
            am[1] = dp4a_uu(W[w], __byte_perm(XL[2 * j], XL[2 * j + 1], sel), am[1]);
        }
    } else {   // Pm == 2: word w covers k = 16w .. 16w+15; field i of a byte: k 16w + {2i, 8+2i, 2i+1, 9+2i}
        const uint2 w2 = lds64(wb + lane * 8);
        const uint32_t W[2] = {w2.x, w2.y};
#pragma unroll
        for (int w = 0; w < 2; ++w)
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                const int a = 2 * w, b = 2 * w + 1, hi = i & 1;
                const uint32_t sel = (i & 2) ? 0x7351u : 0x6240u;
                const uint32_t t = (W[w] >> (2 * i)) & 0x03030303u;
                am[0] = dp4a_us(t, __byte_perm(XH[2 * a + hi], XH[2 * b + hi], sel), am[0]);
                am[1] = dp4a_uu(t, __byte_perm(XL[2 * a + hi], XL[2 * b + hi], sel), am[1]);
            }
    }
    // ---- extra plane (bits above the main plane), at byte 128 * Pm of the block
    if constexpr (Pe == 1) {   // one word: bit i of a byte: k = {2i, 16+2i, 2i+1, 17+2i}
        const uint32_t w = lds32(wb + 128 * Pm + lane * 4);
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            const int oa = i >> 2, ob = 2 + (i >> 2), t4 = i & 3, hi = t4 & 1;
            const uint32_t sel = (t4 & 2) ? 0x7351u : 0x6240u;
            const uint32_t t = (w >> i) & 0x01010101u;
            ae[0] = dp4a_us(t, __byte_perm(XH[2 * oa + hi], XH[2 * ob + hi], sel), ae[0]);
            ae[1] = dp4a_uu(t, __byte_perm(XL[2 * oa + hi], XL[2 * ob + hi], sel), ae[1]);
        }
    } else if constexpr (Pe == 2) {
        const uint2 w2 = lds64(wb + 128 * Pm + lane * 8);
        const uint32_t W[2] = {w2.x, w2.y};
#pragma unroll
        for (int w = 0; w < 2; ++w)
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                const int a = 2 * w, b = 2 * w + 1, hi = i & 1;
                const uint32_t sel = (i & 2) ? 0x7351u : 0x6240u;
                const uint32_t t = (W[w] >> (2 * i)) & 0x03030303u;
                ae[0] = dp4a_us(t, __byte_perm(XH[2 * a + hi], XH[2 * b + hi], sel), ae[0]);
                ae[1] = dp4a_uu(t, __byte_perm(XL[2 * a + hi], XL[2 * b + hi], sel), ae[1]);
            }
    }
}

template <int BITS>
__device__ __forceinline__ int consume_stage(uint32_t slot, int n, uint32_t xs, uint32_t asum, int lane, int (&am)[4], int (&ae)[2]) {
    constexpr uint32_t bb = 128 * BITS;
    int S = 0;
#pragma unroll 2
    for (int s = 0; s < n; ++s) {
        consume_slab<BITS>(slot + s * bb, xs + s * 64, lane, am, ae);
        S += (int)lds32(asum + s * 4);
    }
    return S;
}

// scale / zero-point of one (group, column), fetched a segment ahead of its use
struct RawScale {
    uint32_t w;        // EXL2: q_scale word of the lane's column;  GPTQ: qzeros word
    half hs;           // GPTQ: fp16 scale;  EXL2: q_scale_max[group]
};
__device__ __forceinline__ RawScale load_scales(const I8Params& P, int mi, int blk, int group, int lane) {
    const I8Mat& m = P.mat[mi];
    const int n = (blk - m.blk_base) * 32 + lane;
    RawScale r;
    r.w = 0u;
    r.hs = __ushort_as_half(0);
    if (n < m.N) {
        if (!m.is_gptq) {
            r.w = __ldg(m.q_scale + (size_t)group * (m.N >> 3) + (n >> 3));
            r.hs = __ldg(m.q_scale_max + group);
        } else {
            r.w = __ldg(m.qzeros + (size_t)group * (m.N >> 3) + (n >> 3));
            r.hs = __ldg(m.gptq_scales + (size_t)group * m.N + n);
        }
    }
    return r;
}

__device__ __forceinline__ void finalize_block(const I8Params& P, int blk, int lane, float v) {
    int mi = 0;
#pragma unroll
    for (int i = 1; i < I8_MAX_MATS; ++i)
        if (i < P.num_mats && blk >= P.mat[i].blk_base) mi = i;
    const I8Mat& m = P.mat[mi];
    const int n = (blk - m.blk_base) * 32 + lane;
    if (n < m.N) {
        if (m.bias) v += __half2float(m.bias[n]);
        if (!m.clear) v += __half2float(m.c[n]);
        const half h = __float2half_rn(v);
        m.c[n] = h;
        if (m.c_perm) m.c_perm[m.out_invperm ? (int)m.out_invperm[n] : n] = h;
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

constexpr int DF_FLUSH = 1, DF_BLOCK_DONE = 2;

__device__ __forceinline__ unsigned long long i8_gtimer() {
    unsigned long long t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    return t;
}
// stamps of thread 0 of CTA dbg_cta: 0 start, 1 first stages requested, 2 dependency wait over, 3 row staged,
// 4 warp 0's main loop done, 5 all warps done, (6 = earliest CTA start, 7 = latest CTA end over the grid)
#define I8_STAMP(i) do { if (P.dbg) { if (blockIdx.x == P.dbg_cta && tid == 0) P.dbg[i] = i8_gtimer(); if ((i) == 0 && tid == 0) atomicMin(P.dbg + 6, i8_gtimer()); } } while (0)

// ---- the kernel -----------------------------------------------------------------------------------------------------------
template <int I8_WARPS>
__global__ void __launch_bounds__(I8_WARPS * 32, 2) gemv_i8_kernel(const __grid_constant__ I8Params P) {
    constexpr int I8_THREADS = I8_WARPS * 32;
    extern __shared__ __align__(128) uint8_t smem[];
    __shared__ uint64_t bars[I8_WARPS * I8_MAX_STAGES];
    __shared__ float s_red[I8_WARPS];
    __shared__ int em_blk[I8_WARPS][2], em_n[I8_WARPS][2];

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int KS = P.KS, ns = P.ns;
    I8_STAMP(0);
    if (tid < I8_WARPS * I8_MAX_STAGES) mbar_init(smem_addr(&bars[tid]), 1);
    if (tid < I8_WARPS * 2) em_blk[tid >> 1][tid & 1] = -1;
    mbar_fence_init();
    __syncthreads();
    griddep_launch_dependents();          // the next launch may become resident (and prefetch ITS weights) right away

    // ---- this CTA's blocks / this warp's unit range (unit = (block, slab); CTA-relative linear index b * KS + ks)
    const int blk0 = P.cta_blk[blockIdx.x], nb = (int)P.cta_blk[blockIdx.x + 1] - blk0;
    if (nb <= 0) return;
    const int units = nb * KS;
    const int l0 = (units * warp) / I8_WARPS, l1 = (units * (warp + 1)) / I8_WARPS;

    // shared-memory map: generic pointers for the prologue's stores, 32-bit shared-space addresses (`lds*`) for the main loop
    const I8Smem sm = i8_smem_map(I8_WARPS, ns, KS, P.lcap, P.nb_max);
    uint8_t* const act_g = smem + sm.act;
    int* const asum_s = reinterpret_cast<int*>(smem + sm.asum);
    float* const ascale_s = reinterpret_cast<float*>(smem + sm.ascale);
    float* const emit_base = reinterpret_cast<float*>(smem + sm.emit);
    uint2* const list_g = reinterpret_cast<uint2*>(smem + sm.list) + (size_t)warp * P.lcap;
    unsigned long long* const blksrc_g = reinterpret_cast<unsigned long long*>(smem + sm.blksrc);
    uint32_t sbase;        // kept opaque: the compiler would otherwise re-derive every shared address from S2R in the loop
    asm volatile("mov.u32 %0, %1;" : "=r"(sbase) : "r"(smem_addr(smem)));
    const uint32_t ring = sbase + (uint32_t)warp * (uint32_t)(ns * I8_SLOT_BYTES);
    const uint32_t act = sbase + sm.act, asum = sbase + sm.asum, ascale = sbase + sm.ascale;
    const uint32_t list = sbase + sm.list + (uint32_t)warp * (uint32_t)P.lcap * 8u;
    uint32_t bar0;
    asm volatile("mov.u32 %0, %1;" : "=r"(bar0) : "r"(smem_addr(&bars[warp * I8_MAX_STAGES])));

    // per-block table: byte stream of block b (and its matrix, in the low 2 bits: streams are 16-byte aligned)
    for (int b = tid; b < nb; b += I8_THREADS) {
        const int blk = blk0 + b;
        int mi = 0;
#pragma unroll
        for (int i = 1; i < I8_MAX_MATS; ++i)
            if (i < P.num_mats && blk >= P.mat[i].blk_base) mi = i;
        const I8Mat& m = P.mat[mi];
        blksrc_g[b] = (unsigned long long)(m.packed + (size_t)(blk - m.blk_base) * m.blk_stream_bytes) | (unsigned long long)mi;
    }

    // ---- this warp's STAGE LIST, built here (nothing below depends on the previous launch): one 8-byte descriptor per stage
    //        w0 = byte offset in the block stream | slabs << 22 | bits << 25 | flags << 29      w1 = ks | group << 11 | block << 22
    //      so the main loop does no position arithmetic at all.  A stage is at most 2 KB (4 slabs up to 4 bits, 2 above) and
    //      never crosses a quantisation group, a 128-k row block, a bit-width region or the end of the warp's range.
    int nst = 0;
    {
        int lin = l0, b = l0 / KS, ks = l0 - (l0 / KS) * KS, mi = 0, r = 0;
        auto set_block = [&]() {
            const int blk = blk0 + b;
            mi = 0;
#pragma unroll
            for (int i = 1; i < I8_MAX_MATS; ++i)
                if (i < P.num_mats && blk >= P.mat[i].blk_base) mi = i;
        };
        int r_begin = 0, r_bits = 4, r_spg = 0, r_gbase = 0, r_end = 0;
        uint32_t r_off = 0;
        auto set_region = [&]() {
            const I8Mat& m = P.mat[mi];
            const QRegion& rg = m.reg[r];
            r_begin = rg.ks_begin; r_bits = rg.bits; r_spg = rg.spg_log2; r_gbase = rg.group_base; r_off = rg.off_base;
            r_end = (r + 1 < m.num_regions) ? m.reg[r + 1].ks_begin : KS;
        };
        if (lin < l1) {
            set_block();
            const I8Mat& m = P.mat[mi];
#pragma unroll
            for (int i = 1; i < MAX_REGIONS; ++i)
                if (i < m.num_regions && ks >= m.reg[i].ks_begin) r = i;
            set_region();
        }
        while (lin < l1) {
            const int rel = ks - r_begin, g = rel >> r_spg;
            const int gend = r_begin + ((g + 1) << r_spg);
            const int segend = min(min(gend, (ks | 3) + 1), min(r_end, ks + (l1 - lin)));
            const int n = min(segend - ks, r_bits > 4 ? 2 : 4);
            const uint32_t flags = ((ks + n == segend) ? DF_FLUSH : 0) | ((ks + n == KS || lin + n == l1) ? DF_BLOCK_DONE : 0);
            if (nst >= P.lcap) __trap();
            if (lane == 0)
                list_g[nst] = make_uint2((r_off + (uint32_t)(rel * 128 * r_bits)) | ((uint32_t)n << 22) | ((uint32_t)r_bits << 25) | (flags << 29),
                                         (uint32_t)ks | ((uint32_t)(r_gbase + g) << 11) | ((uint32_t)b << 22));
            ++nst;
            lin += n;
            ks += n;
            if (ks >= KS) {
                ks = 0;
                ++b;
                r = 0;
                if (lin < l1) { set_block(); set_region(); }
            } else if (ks >= r_end) {
                ++r;
                set_region();
            }
        }
    }
    __syncthreads();          // block table + stage lists visible
    auto issue_stage = [&](int s, int slot_idx) {          // lane 0: request stage s into ring slot slot_idx
        const uint2 d = lds64(list + (uint32_t)s * 8u);
        const uint32_t bytes = ((d.x >> 22) & 7u) * 128u * ((d.x >> 25) & 15u);
        const unsigned long long src = *reinterpret_cast<const volatile unsigned long long*>(blksrc_g + (d.y >> 22));
        const uint32_t bar = bar0 + (uint32_t)slot_idx * 8u;
        mbar_arrive_expect_tx(bar, bytes);
        bulk_copy_g2s(ring + (uint32_t)slot_idx * I8_SLOT_BYTES, reinterpret_cast<const uint8_t*>(src & ~3ull) + (d.x & 0x3fffffu), bytes, bar);
    };
    if (lane == 0)
        for (int i = 0; i < ns && i < nst; ++i) issue_stage(i, i);
    RawScale raw = {0u, __ushort_as_half(0)};
    if (nst > 0) {
        const uint2 d0 = lds64(list);
        const unsigned long long src = *reinterpret_cast<const volatile unsigned long long*>(blksrc_g + (d0.y >> 22));
        raw = load_scales(P, (int)(src & 3ull), blk0 + (int)(d0.y >> 22), (int)((d0.y >> 11) & 0x7ffu), lane);
    }

    // ---- static operands of the prologue, fetched before the dependency wait: permutation indices (when the row has to be
    //      gathered, or the norm weight has) and the norm weight of this thread's first octets
    const int n_oct = P.K >> 3;
    const bool gather_x = P.perm != nullptr && !P.x_permuted;
    constexpr int PF = 3;
    uint4 pv[PF], wv[PF];
#pragma unroll
    for (int r = 0; r < PF; ++r) {
        const int o = r * I8_THREADS + tid;
        pv[r] = make_uint4(0, 0, 0, 0);
        wv[r] = make_uint4(0, 0, 0, 0);
        if (o < n_oct) {
            if (P.perm) pv[r] = __ldg(reinterpret_cast<const uint4*>(P.perm + o * 8));
            if (P.mode == I8_RMSNORM) {
                if (P.perm) {
                    const uint16_t* pi = reinterpret_cast<const uint16_t*>(&pv[r]);
                    uint16_t* wo = reinterpret_cast<uint16_t*>(&wv[r]);
#pragma unroll
                    for (int e = 0; e < 8; ++e) wo[e] = __half_as_ushort(__ldg(P.norm_w + pi[e]));
                } else {
                    wv[r] = __ldg(reinterpret_cast<const uint4*>(P.norm_w + o * 8));
                }
            }
        }
    }

    I8_STAMP(1);
    griddep_wait();                        // everything below may read what the previous launch wrote
    I8_STAMP(2);

    // ---- prologue: the row -> (optional RMSNorm weight / act*mul) -> 16-bit integers per 128-k block -> shared memory.
    //      Every CTA stages the whole row (its blocks span all of K); 1/rms is applied to the finished fp32 sums.
    float sumsq = 0.f;
    // A row that has to be gathered through q_perm is first copied into shared memory with coalesced 16-byte loads (into the
    // region that will hold the staged integers: same size) and gathered from there: one L2 round trip instead of eight
    // scattered 2-byte sector reads per thread.
    const bool two_in = (P.mode == I8_SILU_MUL || P.mode == I8_GELU_MUL);
    const bool smem_gather = gather_x && !two_in && n_oct <= PF * I8_THREADS;
    uint4 hg[PF];
    if (smem_gather) {
        for (int o = tid; o < n_oct; o += I8_THREADS)
            reinterpret_cast<uint4*>(act_g)[o] = __ldcg(reinterpret_cast<const uint4*>(P.x + o * 8));
        __syncthreads();
#pragma unroll
        for (int r = 0; r < PF; ++r) {
            const int o = r * I8_THREADS + tid;
            hg[r] = make_uint4(0, 0, 0, 0);
            if (o < n_oct) {
                const uint16_t* pi = reinterpret_cast<const uint16_t*>(&pv[r]);
                uint16_t* ho = reinterpret_cast<uint16_t*>(&hg[r]);
#pragma unroll
                for (int e = 0; e < 8; ++e) ho[e] = reinterpret_cast<const uint16_t*>(act_g)[pi[e]];
            }
        }
        __syncthreads();
    }
    auto stage_round = [&](int o, uint4 pidx, uint4 wreg, bool pre, uint4 hval) {
        const bool valid = o < n_oct;
        float f[8];
#pragma unroll
        for (int e = 0; e < 8; ++e) f[e] = 0.f;
        if (valid) {
            const uint16_t* pi = reinterpret_cast<const uint16_t*>(&pidx);
            half h[8], h2[8];
            const bool two = two_in;
            if (pre) {
                *reinterpret_cast<uint4*>(h) = hval;
            } else if (gather_x) {
#pragma unroll
                for (int e = 0; e < 8; ++e) h[e] = __ldcg(P.x + pi[e]);
                if (two) {
#pragma unroll
                    for (int e = 0; e < 8; ++e) h2[e] = __ldcg(P.x2 + pi[e]);
                }
            } else {
                const uint4 v = __ldcg(reinterpret_cast<const uint4*>(P.x + o * 8));
                *reinterpret_cast<uint4*>(h) = v;
                if (two) *reinterpret_cast<uint4*>(h2) = __ldcg(reinterpret_cast<const uint4*>(P.x2 + o * 8));
            }
            if (P.mode == I8_RMSNORM) {
                const half* wh = reinterpret_cast<const half*>(&wreg);
#pragma unroll
                for (int e = 0; e < 8; ++e) {
                    const float xf = fmaxf(-65504.f, fminf(__half2float(h[e]), 65504.f));
                    sumsq = fmaf(xf, xf, sumsq);
                    f[e] = xf * __half2float(wh[e]);
                }
            } else if (P.mode == I8_SILU_MUL) {
#pragma unroll
                for (int e = 0; e < 8; ++e) f[e] = __half2float(__hmul(silu_h(h[e]), h2[e]));
            } else if (P.mode == I8_GELU_MUL) {
#pragma unroll
                for (int e = 0; e < 8; ++e) f[e] = __half2float(__hmul(gelu_h(h[e]), h2[e]));
            } else {
#pragma unroll
                for (int e = 0; e < 8; ++e) f[e] = __half2float(h[e]);
            }
        }
        float amax = 0.f;
#pragma unroll
        for (int e = 0; e < 8; ++e) amax = fmaxf(amax, fabsf(f[e]));
#pragma unroll
        for (int s = 1; s < 16; s <<= 1) amax = fmaxf(amax, __shfl_xor_sync(0xffffffffu, amax, s));      // 16 lanes = one 128-k block
        // power-of-two block scale 2^e with max * 2^e in [2^14, 2^15): fp16 row values within a factor 16 of the block maximum
        // are represented EXACTLY and all scale products are exact -- a unit-vector row returns reconstruct()'s fp16 weights
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
            const int si = o >> 2, j = o & 3;
            uint2 hw, lw;
            hw.x = pack(0, 4, 1, 5, 8); hw.y = pack(2, 6, 3, 7, 8);
            lw.x = pack(0, 4, 1, 5, 0); lw.y = pack(2, 6, 3, 7, 0);
            *reinterpret_cast<uint2*>(act_g + (size_t)si * 64 + j * 8) = hw;
            *reinterpret_cast<uint2*>(act_g + (size_t)si * 64 + 32 + j * 8) = lw;
            if (j == 0) asum_s[si] = sum;
            if ((o & 15) == 0) ascale_s[si >> 2] = amax > 0.f ? __uint_as_float((ef - 14u) << 23) : 0.f;
        }
    };
#pragma unroll
    for (int r = 0; r < PF; ++r)
        if (r * I8_THREADS < n_oct) stage_round(r * I8_THREADS + tid, pv[r], wv[r], smem_gather, hg[r]);
    for (int ob = PF * I8_THREADS; ob < n_oct; ob += I8_THREADS) {
        const int o = ob + tid;
        uint4 pidx = make_uint4(0, 0, 0, 0), wreg = make_uint4(0, 0, 0, 0);
        if (o < n_oct) {
            if (P.perm) pidx = __ldg(reinterpret_cast<const uint4*>(P.perm + o * 8));
            if (P.mode == I8_RMSNORM) {
                if (P.perm) {
                    const uint16_t* pi = reinterpret_cast<const uint16_t*>(&pidx);
                    uint16_t* wo = reinterpret_cast<uint16_t*>(&wreg);
#pragma unroll
                    for (int e = 0; e < 8; ++e) wo[e] = __half_as_ushort(__ldg(P.norm_w + pi[e]));
                } else {
                    wreg = __ldg(reinterpret_cast<const uint4*>(P.norm_w + o * 8));
                }
            }
        }
        stage_round(o, pidx, wreg, false, make_uint4(0, 0, 0, 0));
    }
    if (P.mode == I8_RMSNORM) {
#pragma unroll
        for (int s = 16; s > 0; s >>= 1) sumsq += __shfl_xor_sync(0xffffffffu, sumsq, s);
        if (lane == 0) s_red[warp] = sumsq;
    }
    __syncthreads();
    I8_STAMP(3);
    float rrms = 1.f;
    if (P.mode == I8_RMSNORM) {
        float t = 0.f;
#pragma unroll
        for (int i = 0; i < I8_WARPS; ++i) t += s_red[i];
        rrms = rsqrtf(t * (1.0f / (float)P.K) + P.norm_eps);
    }

    // ---- main loop: this warp alone, stage by stage; everything positional comes from the stage list.
    int am[4] = {0, 0, 0, 0}, ae[2] = {0, 0};
    float tot = 0.f;
    int S = 0, cslot = 0, blk_slabs = 0, emits = 0;
    uint32_t phase = 0;
    for (int s = 0; s < nst; ++s) {
        mbar_wait(bar0 + (uint32_t)cslot * 8u, (phase >> cslot) & 1u);
        phase ^= 1u << cslot;
        const uint2 d = lds64(list + (uint32_t)s * 8u);
        const int n = (int)((d.x >> 22) & 7u), bits = (int)((d.x >> 25) & 15u), ks = (int)(d.y & 0x7ffu);
        const uint32_t slot = ring + (uint32_t)cslot * I8_SLOT_BYTES;
        const uint32_t xs = act + (uint32_t)ks * 64u, as = asum + (uint32_t)ks * 4u;
        switch (bits) {
            case 4: S += consume_stage<4>(slot, n, xs, as, lane, am, ae); break;
            case 5: S += consume_stage<5>(slot, n, xs, as, lane, am, ae); break;
            case 6: S += consume_stage<6>(slot, n, xs, as, lane, am, ae); break;
            case 3: S += consume_stage<3>(slot, n, xs, as, lane, am, ae); break;
            case 8: S += consume_stage<8>(slot, n, xs, as, lane, am, ae); break;
            default: S += consume_stage<2>(slot, n, xs, as, lane, am, ae); break;
        }
        __syncwarp();
        if (lane == 0 && s + ns < nst) issue_stage(s + ns, cslot);            // refill the slot just drained
        cslot = (cslot + 1 == ns) ? 0 : cslot + 1;
        blk_slabs += n;
        if (d.x & ((uint32_t)DF_FLUSH << 29)) {
            // integer sums -> fp32:  sum_k a_k (q_k - zero) * scale  =  (sum a q - zero * sum a) * scale_w * scale_row
            const int b = (int)(d.y >> 22);
            const unsigned long long src = *reinterpret_cast<const volatile unsigned long long*>(blksrc_g + b);
            const I8Mat& m = P.mat[(int)(src & 3ull)];
            const int col = (blk0 + b - m.blk_base) * 32 + lane;
            int v = ((am[0] << 8) + am[1]) + (((am[2] << 8) + am[3]) >> 4) + (((ae[0] << 8) + ae[1]) << plane_main(bits));
            const uint32_t nib = (raw.w >> ((col & 7) * 4)) & 15u;
            float ws;
            int zero;
            if (!m.is_gptq) {
                const int qs = (int)nib + 1;
                ws = __half2float(__hmul(__int2half_rn(qs * qs), raw.hs));          // dq_scale, cuda/quant/qdq_util.cuh:24-30
                zero = 1 << (bits - 1);
            } else {
                ws = __half2float(raw.hs);
                zero = (int)nib + 1;                                                 // q_gemm_kernel_gptq.cuh:167-172
            }
            v -= zero * S;
            tot = fmaf((float)v, ws * __uint_as_float(lds32(ascale + (uint32_t)(ks >> 2) * 4u)), tot);
            am[0] = am[1] = am[2] = am[3] = 0;
            ae[0] = ae[1] = 0;
            S = 0;
            if (s + 1 < nst) {                 // scales of the next segment
                const uint2 dn = lds64(list + (uint32_t)(s + 1) * 8u);
                const unsigned long long sn = *reinterpret_cast<const volatile unsigned long long*>(blksrc_g + (dn.y >> 22));
                raw = load_scales(P, (int)(sn & 3ull), blk0 + (int)(dn.y >> 22), (int)((dn.y >> 11) & 0x7ffu), lane);
            }
            if (d.x & ((uint32_t)DF_BLOCK_DONE << 29)) {
                if (blk_slabs == KS) {
                    finalize_block(P, blk0 + b, lane, tot * rrms);          // this warp covered the block's whole K by itself
                } else {
                    if (emits >= 2) __trap();       // a warp's range has at most two partial blocks (its first and its last)
                    emit_base[(warp * 2 + emits) * 32 + lane] = tot;
                    if (lane == 0) { em_blk[warp][emits] = blk0 + b; em_n[warp][emits] = blk_slabs; }
                    ++emits;
                }
                tot = 0.f;
                blk_slabs = 0;
            }
        }
    }
    I8_STAMP(4);
    __syncthreads();
    I8_STAMP(5);

    // ---- split-K never left the CTA: sum the warps' partials of each block in warp order, finalise
    for (int b = warp; b < nb; b += I8_WARPS) {
        const int blk = blk0 + b;
        float v = 0.f;
        int cnt = 0;
#pragma unroll
        for (int w = 0; w < I8_WARPS; ++w)
#pragma unroll
            for (int e = 0; e < 2; ++e)
                if (em_blk[w][e] == blk) {
                    v += emit_base[(w * 2 + e) * 32 + lane];
                    cnt += em_n[w][e];
                }
        if (cnt == 0) continue;          // finalised by the one warp that covered it
        if (cnt != KS) __trap();
        finalize_block(P, blk, lane, v * rrms);
    }
    if (P.dbg && lane == 0) atomicMax(P.dbg + 7, i8_gtimer());
}

// ---- host side ---------------------------------------------------------------------------------------------------------------

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

// Contiguous partition of the launch's blocks over at most `ctas` CTAs minimising the largest CTA (in bytes).
void i8_partition_blocks(const std::vector<uint32_t>& bytes, int ctas, unsigned short* out, int* used) {
    const int nb = (int)bytes.size();
    unsigned long long total = 0, biggest = 0;
    for (uint32_t b : bytes) { total += b; biggest = std::max<unsigned long long>(biggest, b); }
    unsigned long long lo = std::max(biggest, (total + ctas - 1) / ctas), hi = total;
    auto fits = [&](unsigned long long cap) {
        int c = 1;
        unsigned long long acc = 0;
        for (uint32_t b : bytes) {
            if (acc + b > cap) { ++c; acc = 0; }
            acc += b;
        }
        return c <= ctas;
    };
    while (lo < hi) {
        const unsigned long long mid = (lo + hi) / 2;
        if (fits(mid)) hi = mid; else lo = mid + 1;
    }
    int c = 0;
    unsigned long long acc = 0;
    out[0] = 0;
    for (int b = 0; b < nb; ++b) {
        if (acc + bytes[b] > lo) { out[++c] = (unsigned short)b; acc = 0; }
        acc += bytes[b];
    }
    out[++c] = (unsigned short)nb;
    *used = c;
}

int gemv_i8_launch(int device, cudaStream_t stream, const I8Out* outs, int nm, const I8Input& in) {
    EXL2B_REQUIRE(nm >= 1 && nm <= I8_MAX_MATS, "bad matrix count %d", nm);
    EXL2B_REQUIRE(in.x, "null input row");
    EXL2B_REQUIRE(in.mode != I8_RMSNORM || in.norm_w, "RMSNorm prologue without a weight");
    EXL2B_REQUIRE((in.mode != I8_SILU_MUL && in.mode != I8_GELU_MUL) || in.x2, "act*mul prologue without the second operand");
    // warps per CTA: 16 (64 registers / thread) or 12 (80) -- EXL2B_I8_WARPS selects, both keep two CTAs per SM resident
    static const int warps = [] {
        const char* e = getenv("EXL2B_I8_WARPS");
        const int w = e ? atoi(e) : 16;
        return (w == 12 || w == 8) ? w : 16;
    }();
    static bool attr_set[64] = {false};
    if (device >= 0 && device < 64 && !attr_set[device]) {
        EXL2B_CUDA(cudaFuncSetAttribute(gemv_i8_kernel<16>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
        EXL2B_CUDA(cudaFuncSetAttribute(gemv_i8_kernel<12>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
        EXL2B_CUDA(cudaFuncSetAttribute(gemv_i8_kernel<8>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
        attr_set[device] = true;
    }

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
    P.x_permuted = in.x_permuted;
    std::vector<uint32_t> blk_bytes;
    int blk = 0;
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
        m.c_perm = outs[i].c_perm;
        m.out_invperm = outs[i].out_invperm;
        m.clear = outs[i].clear;
        m.blk_stream_bytes = v.blk_stream_bytes;
        m.N = v.N;
        m.blk_base = blk;
        m.is_gptq = v.is_gptq;
        m.num_regions = v.num_regions;
        for (int r = 0; r < v.num_regions; ++r) m.reg[r] = v.reg[r];
        // only blocks that hold real columns (the last strip of a padded matrix may contain all-padding blocks)
        const int nblk = (v.N + 31) / 32;
        for (int b = 0; b < nblk; ++b) blk_bytes.push_back(v.blk_stream_bytes);
        blk += nblk;
    }
    EXL2B_REQUIRE(blk < 65535, "too many column blocks (%d)", blk);
    EXL2B_REQUIRE((long long)blk * P.KS < (1ll << 26), "problem too large for 32-bit unit arithmetic");
    const int sms = std::min(device_sm_count(device), I8_MAX_CTAS);
    int C = 0;
    i8_partition_blocks(blk_bytes, sms, P.cta_blk, &C);

    // shared memory: i8_smem_map (weight rings, staged row, per-warp partials, stage lists, block table)
    // stage-list capacity: an upper bound of the stages one warp can have (every boundary a stage may not cross adds at most one)
    int nb_max = 0, smin = 4, max_regions = 1;
    for (int c = 0; c < C; ++c) nb_max = std::max(nb_max, (int)P.cta_blk[c + 1] - (int)P.cta_blk[c]);
    for (int i = 0; i < nm; ++i) {
        max_regions = std::max(max_regions, P.mat[i].num_regions);
        for (int r = 0; r < P.mat[i].num_regions; ++r) {
            smin = std::min(smin, std::min(1 << P.mat[i].reg[r].spg_log2, P.mat[i].reg[r].bits > 4 ? 2 : 4));
            if (P.mat[i].reg[r].ks_begin & 3) smin = 1;      // groups not aligned with the 128-k row blocks: segments may be single slabs
        }
    }
    const int upw = (nb_max * P.KS + warps - 1) / warps + 1;
    P.lcap = (upw + smin - 1) / smin + (upw / P.KS + 2) * (max_regions + 1) + 4;
    P.nb_max = nb_max;
    EXL2B_REQUIRE(P.KS <= 2048 && nb_max < 1024, "matrix too large for the stage descriptor (K <= 65536)");
    auto smem_for = [&](int ns) { return (size_t)i8_smem_map(warps, ns, P.KS, P.lcap, P.nb_max).total; };
    // two launches co-resident per SM (227 KB, 1 KB reserved per CTA) is what lets the next launch prefetch: the deepest ring
    // (<= 4 slots per warp) that keeps the CTA <= 112 KB, never fewer than 2 slots
    P.ns = I8_MAX_STAGES;
    while (P.ns > 2 && smem_for(P.ns) > 112 * 1024) --P.ns;
    static const int ns_override = [] { const char* e = getenv("EXL2B_I8_NS"); return e ? atoi(e) : 0; }();     // tuning knob
    if (ns_override >= 2 && ns_override <= I8_MAX_STAGES) P.ns = ns_override;
    const size_t smem_total = smem_for(P.ns);
    EXL2B_REQUIRE(smem_total <= 200 * 1024, "shared memory budget exceeded (%zu bytes, K = %d)", smem_total, P.K);
    extern unsigned long long* g_dbg;
    extern int g_dbg_cta, g_dbg_slot;
    P.dbg = g_dbg ? g_dbg + 32 * (g_dbg_slot++ % 64) : nullptr;
    P.dbg_cta = g_dbg_cta;
    if (warps == 16) EXL2B_CUDA(launch_pdl(gemv_i8_kernel<16>, dim3(C), dim3(16 * 32), smem_total, stream, P));
    else if (warps == 12) EXL2B_CUDA(launch_pdl(gemv_i8_kernel<12>, dim3(C), dim3(12 * 32), smem_total, stream, P));
    else EXL2B_CUDA(launch_pdl(gemv_i8_kernel<8>, dim3(C), dim3(8 * 32), smem_total, stream, P));
    return 0;
}

}  // namespace exl2b

// host-only diagnostics hook (tests/test_i8_emulation.py): the block -> CTA partition gemv_i8_launch would use
extern "C" int exl2b_debug_partition(const uint32_t* block_bytes, int num_blocks, int ctas, uint16_t* out, int* used) {
    EXL2B_REQUIRE(block_bytes && out && used && num_blocks > 0 && ctas > 0 && ctas <= exl2b::I8_MAX_CTAS, "bad argument");
    std::vector<uint32_t> b(block_bytes, block_bytes + num_blocks);
    exl2b::i8_partition_blocks(b, ctas, out, used);
    return 0;
}
