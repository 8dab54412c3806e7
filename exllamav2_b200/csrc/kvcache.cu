// Q4 K/V cache pack / unpack (exllamav2_ext/cuda/cache_q.cuh, cuda/cache.cu:143-497).
//
// Same arithmetic as the reference, op for op (so results are bit-identical to it on the same inputs):
//   pack:   warp butterfly Hadamard-32 on two interleaved 32-vectors (unnormalised, fp16 adds), absmax over 32
//           consecutive values, w = w / absmax * 8 + 8, q = clamp(rn(w), 0, 15), scale = absmax / 8
//   unpack: (q - 8) * scale -> Hadamard -> * 1/32
// What differs is the mapping to the machine: the reference runs 256-thread CTAs over 512-value blocks staged
// through shared memory and, for the paged form, a (pages, 256, 2*batch) grid of mostly-empty CTAs
// (cache.cu:244-248).  The math is per 64-value warp unit, so here every warp independently converts 64-value
// units with direct coalesced global access (128 B in / 32 B + 4 B out per unit), the unit list is flattened
// over (k|v, sequence, token range) and the grid is sized to the actual work.
#include <algorithm>

#include "common.cuh"

namespace exl2b {

__device__ __forceinline__ half2 hadamard32(half2 w2, int lane) {
#pragma unroll
    for (int i = 1; i < 32; i <<= 1) {
        const half2 pw2 = __shfl_xor_sync(0xffffffffu, w2, i);
        uint32_t* w2i = reinterpret_cast<uint32_t*>(&w2);
        const int32_t sfm = -static_cast<int32_t>(lane & i) >> 31;
        *w2i ^= (sfm & 0x80008000);
        w2 = __hadd2(w2, pw2);
    }
    return w2;
}

// one warp: 64 fp16 values at `in` -> 32 packed bytes at `out`, 2 scales at `scales`
__device__ __forceinline__ void pack_unit_q4(const half* __restrict__ in, uint8_t* __restrict__ out,
                                             half* __restrict__ scales, int lane) {
    half2 w2 = reinterpret_cast<const half2*>(in)[lane];
    w2 = hadamard32(w2, lane);
    half2 absmax2 = __habs2(w2);
    half absmax = __hmax(__low2half(absmax2), __high2half(absmax2));
    absmax = __hmax(absmax, __shfl_xor_sync(0xffffffffu, absmax, 8));
    absmax = __hmax(absmax, __shfl_xor_sync(0xffffffffu, absmax, 4));
    absmax = __hmax(absmax, __shfl_xor_sync(0xffffffffu, absmax, 2));
    absmax = __hmax(absmax, __shfl_xor_sync(0xffffffffu, absmax, 1));
    absmax2 = __half2half2(absmax);
    const half2 c_8 = __half2half2(__float2half_rn(8));
    const half c_i = __float2half_rn(1.0f / 8.0f);
    w2 = __h2div(w2, absmax2);
    w2 = __hfma2(w2, c_8, c_8);
    const int q0 = min(max(__half2int_rn(__low2half(w2)), 0), 15);
    const int q1 = min(max(__half2int_rn(__high2half(w2)), 0), 15);
    uint32_t q = q0 | (q1 << 4);
    q |= (__shfl_down_sync(0xffffffffu, q, 1) << 8);
    q |= (__shfl_down_sync(0xffffffffu, q, 2) << 16);
    if ((lane & 3) == 0) reinterpret_cast<uint32_t*>(out)[lane >> 2] = q;
    if ((lane & 15) == 0) scales[lane >> 4] = __hmul(absmax, c_i);
}

__device__ __forceinline__ void unpack_unit_q4(const uint8_t* __restrict__ in, const half* __restrict__ scales,
                                               half* __restrict__ out, int lane) {
    const half scale = __ldg(scales + (lane >> 4));
    const uint32_t q = __ldg(reinterpret_cast<const uint32_t*>(in) + (lane >> 2));
    const int shift0 = (lane & 3) * 8;
    const int q0 = ((int)((q >> shift0) & 0x0f)) - 8;
    const int q1 = ((int)((q >> (shift0 + 4)) & 0x0f)) - 8;
    half2 w2 = __halves2half2(__int2half_rn(q0), __int2half_rn(q1));
    w2 = __hmul2(w2, __half2half2(scale));
    w2 = hadamard32(w2, lane);
    w2 = __hmul2(w2, __float2half2_rn(1.0f / 32.0f));
    __stcg(reinterpret_cast<half2*>(out) + lane, w2);
}

struct KvJob {
    const void* k_a; void* k_b; void* k_s;     // pack: fp16 in, u8 out, scales out;  unpack: u8 in, fp16 out, scales in
    const void* v_a; void* v_b; void* v_s;
    int batch, dim, seq_stride;                // non-paged
    int offset_el, width_el;                   // non-paged element range per batch row (multiples of 64)
    int page_size, pages_per_seq, q_len;       // paged
    const int32_t* cache_seqlens;
    const int32_t* block_table;
    int units_per_seq;                         // upper bound of 64-value units per (k|v, sequence)
    int pack;
};

// Resolve unit index -> element offset.  Returns false when the unit is outside the sequence's live range.
__device__ __forceinline__ bool kv_unit_offset(const KvJob& J, int seq, int unit, size_t& el) {
    if (J.page_size == 0) {
        const int e = J.offset_el + unit * 64;
        if (e >= J.offset_el + J.width_el) return false;
        el = (size_t)seq * J.seq_stride + e;
        return true;
    }
    // paged: token range [a, b) of this sequence, widened to whole 512-value blocks like the reference
    // (cache.cu:177-184 pack, :350-357 unpack) when dim is not a multiple of 512
    const int seqlen = J.cache_seqlens[seq];
    int a, b;
    if (J.pack) { a = seqlen; b = seqlen + J.q_len; }
    else { if (!seqlen) return false; a = 0; b = seqlen; }
    long ea = (long)a * J.dim, eb = (long)b * J.dim;
    if (J.dim % 512) { ea = ea / 512 * 512; eb = (eb + 511) / 512 * 512; }
    const long e = ea + (long)unit * 64;
    if (e >= eb) return false;
    const long tok = e / J.dim;
    const int page_idx = (int)(tok / J.page_size);
    if (page_idx >= J.pages_per_seq) return false;
    const int page = J.block_table[(size_t)seq * J.pages_per_seq + page_idx];
    el = ((size_t)page * J.page_size + (size_t)(tok - (long)page_idx * J.page_size)) * J.dim + (size_t)(e - tok * J.dim);
    return true;
}

__global__ void __launch_bounds__(256) kv_q4_kernel(const __grid_constant__ KvJob J) {
    griddep_launch_dependents();
    griddep_wait();
    const int lane = threadIdx.x & 31;
    const long per_kv = (long)J.batch * J.units_per_seq;
    const int nkv = J.v_a ? 2 : 1;
    const long total = per_kv * nkv;
    const long nwarps = ((long)gridDim.x * blockDim.x) >> 5;
    // grid-stride over 64-value units: the grid is capped at a few CTAs per SM, dead units cost one compare
    for (long wid = ((long)blockIdx.x * blockDim.x + threadIdx.x) >> 5; wid < total; wid += nwarps) {
        const int kv = (int)(wid / per_kv);
        const long r = wid - (long)kv * per_kv;
        const int seq = (int)(r / J.units_per_seq), unit = (int)(r - (long)seq * J.units_per_seq);
        size_t el;
        if (!kv_unit_offset(J, seq, unit, el)) continue;
        const void* a = kv ? J.v_a : J.k_a;
        void* b = kv ? J.v_b : J.k_b;
        void* s = kv ? J.v_s : J.k_s;
        if (J.pack)
            pack_unit_q4((const half*)a + el, (uint8_t*)b + el / 2, (half*)s + el / 32, lane);
        else
            unpack_unit_q4((const uint8_t*)a + el / 2, (const half*)s + el / 32, (half*)b + el, lane);
    }
}

static int kv_launch(KvJob& J, cudaStream_t stream) {
    const long warps = (long)J.batch * J.units_per_seq * (J.v_a ? 2 : 1);
    if (warps <= 0) return 0;
    int dev = 0;
    cudaGetDevice(&dev);
    const long cap = (long)device_sm_count(dev) * 8;
    const long blocks = std::min((warps + 7) / 8, cap);
    EXL2B_CUDA(launch_pdl(kv_q4_kernel, dim3((unsigned)blocks), dim3(256), 0, stream, J));
    return 0;
}

static int kv_common(KvJob& J, int batch, int dim, int seq_stride, int offset, int width, int page_size,
                     const int32_t* cache_seqlens, const int32_t* block_table, int pages_per_seq, int wbits) {
    EXL2B_REQUIRE(wbits == 4, "only the Q4 cache (wbits=4) is implemented; got %d", wbits);
    EXL2B_REQUIRE(dim > 0 && dim % 64 == 0, "kv dim %d must be a multiple of 64", dim);
    J.batch = batch;
    J.dim = dim;
    J.seq_stride = seq_stride;
    J.page_size = page_size;
    J.pages_per_seq = pages_per_seq;
    J.cache_seqlens = cache_seqlens;
    J.block_table = block_table;
    if (page_size == 0) {
        // ext_cache.cpp:150-157: widen [offset, offset+width) tokens to 512-element block boundaries
        if (dim % 512) {
            while (((long)offset * dim) % 512) offset--;
            while (((long)width * dim) % 512) width++;
        }
        J.offset_el = offset * dim;
        J.width_el = width * dim;
        J.units_per_seq = J.width_el / 64;
    } else {
        EXL2B_REQUIRE(cache_seqlens && block_table && pages_per_seq > 0, "paged kv needs cache_seqlens and block_table");
        J.q_len = width;
        const long tokens = J.pack ? (long)width : (long)pages_per_seq * page_size;
        J.units_per_seq = (int)((tokens * dim + 511) / 512 * 8 + 8);
    }
    return 0;
}

}  // namespace exl2b

using namespace exl2b;

extern "C" int exl2b_fp16_to_q_kv(const uint16_t* k_in, uint8_t* k_out, uint16_t* k_scales, const uint16_t* v_in,
                                  uint8_t* v_out, uint16_t* v_scales, int batch, int dim, int seq_stride, int offset,
                                  int width, int page_size, const int32_t* cache_seqlens, const int32_t* block_table,
                                  int pages_per_seq, int wbits, exl2b_stream_t stream) {
    EXL2B_REQUIRE(k_in && k_out && k_scales, "null k tensors");
    KvJob J = {};
    J.pack = 1;
    J.k_a = k_in; J.k_b = k_out; J.k_s = k_scales;
    J.v_a = v_in; J.v_b = v_out; J.v_s = v_scales;
    int rc = kv_common(J, batch, dim, seq_stride, offset, width, page_size, cache_seqlens, block_table, pages_per_seq, wbits);
    if (rc) return rc;
    return kv_launch(J, (cudaStream_t)stream);
}

extern "C" int exl2b_q_to_fp16_kv(const uint8_t* k_in, const uint16_t* k_scales, uint16_t* k_out, const uint8_t* v_in,
                                  const uint16_t* v_scales, uint16_t* v_out, int batch, int dim, int seq_stride,
                                  int offset, int width, int page_size, const int32_t* cache_seqlens,
                                  const int32_t* block_table, int pages_per_seq, int wbits, exl2b_stream_t stream) {
    EXL2B_REQUIRE(k_in && k_out && k_scales, "null k tensors");
    KvJob J = {};
    J.pack = 0;
    J.k_a = k_in; J.k_b = k_out; J.k_s = (void*)k_scales;
    J.v_a = v_in; J.v_b = v_out; J.v_s = (void*)v_scales;
    int rc = kv_common(J, batch, dim, seq_stride, offset, width, page_size, cache_seqlens, block_table, pages_per_seq, wbits);
    if (rc) return rc;
    return kv_launch(J, (cudaStream_t)stream);
}
