// Paged decode attention over the fp16 K/V view -- the stand-in for flash_attn_with_kvcache, which the reference
// calls between q_attn_forward_1 and q_attn_forward_2 (exllamav2/attn.py:602-613; third-party there).  Same
// semantics: the q_len new K/V rows are appended to the paged cache at [seqlen, seqlen + q_len) and the queries
// attend causally over [0, seqlen + i].  Written for the decode regime (q_len <= 8, batch small): one CTA per
// (head, sequence), 4 warps split the context, online softmax in fp32, one pass over K and V.
#include "common.cuh"

namespace exl2b {

struct AttnParams {
    const half* q;          // [batch, q_len, H, hd]
    const half* k_new;      // [batch, q_len, KVH, hd]
    const half* v_new;
    half* k_cache;          // [pages, page_size, KVH, hd]
    half* v_cache;
    const int32_t* cache_seqlens;   // [batch]
    const int32_t* block_table;     // [batch, pages_per_seq]
    half* out;              // [batch, q_len, H, hd]
    int q_len, H, KVH, hd, page_size, pages_per_seq;
    float scale_log2;       // softmax_scale * log2(e)
};

template <int VEC>   // elements per lane: hd = 32 * VEC
__global__ void __launch_bounds__(128) attn_decode_kernel(const __grid_constant__ AttnParams P) {
    griddep_launch_dependents();
    griddep_wait();
    const int h = blockIdx.x, b = blockIdx.y;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int group = P.H / P.KVH, kvh = h / group;
    const int seqlen = P.cache_seqlens[b];
    const int hd = P.hd;
    __shared__ float s_m[4], s_l[4];
    __shared__ float s_acc[4][32 * VEC];

    // append the new rows (one designated head per kv head does it; readers take new rows from k_new / v_new)
    if (h % group == 0) {
        for (int idx = threadIdx.x; idx < P.q_len * hd; idx += blockDim.x) {
            const int i = idx / hd, d = idx - i * hd;
            const int pos = seqlen + i;
            const int page = P.block_table[(size_t)b * P.pages_per_seq + pos / P.page_size];
            const size_t dst = (((size_t)page * P.page_size + pos % P.page_size) * P.KVH + kvh) * hd + d;
            const size_t src = (((size_t)b * P.q_len + i) * P.KVH + kvh) * hd + d;
            P.k_cache[dst] = P.k_new[src];
            P.v_cache[dst] = P.v_new[src];
        }
    }

    for (int i = 0; i < P.q_len; ++i) {
        float qv[VEC];
        const half* qp = P.q + (((size_t)b * P.q_len + i) * P.H + h) * hd + lane * VEC;
#pragma unroll
        for (int j = 0; j < VEC; ++j) qv[j] = __half2float(qp[j]) * P.scale_log2;
        float m = -INFINITY, l = 0.f, acc[VEC];
#pragma unroll
        for (int j = 0; j < VEC; ++j) acc[j] = 0.f;
        const int n_ctx = seqlen + i + 1;
        for (int p = warp; p < n_ctx; p += 4) {
            const half *kp, *vp;
            if (p < seqlen) {
                const int page = P.block_table[(size_t)b * P.pages_per_seq + p / P.page_size];
                const size_t off = (((size_t)page * P.page_size + p % P.page_size) * P.KVH + kvh) * hd + lane * VEC;
                kp = P.k_cache + off;
                vp = P.v_cache + off;
            } else {
                const size_t off = (((size_t)b * P.q_len + (p - seqlen)) * P.KVH + kvh) * hd + lane * VEC;
                kp = P.k_new + off;
                vp = P.v_new + off;
            }
            float s = 0.f, vv[VEC];
#pragma unroll
            for (int j = 0; j < VEC; ++j) {
                s = fmaf(qv[j], __half2float(kp[j]), s);
                vv[j] = __half2float(vp[j]);
            }
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
            const float m_new = fmaxf(m, s);
            const float corr = exp2f(m - m_new), pexp = exp2f(s - m_new);
            l = l * corr + pexp;
#pragma unroll
            for (int j = 0; j < VEC; ++j) acc[j] = acc[j] * corr + pexp * vv[j];
            m = m_new;
        }
        if (lane == 0) { s_m[warp] = m; s_l[warp] = l; }
#pragma unroll
        for (int j = 0; j < VEC; ++j) s_acc[warp][lane * VEC + j] = acc[j];
        __syncthreads();
        if (warp == 0) {
            float mm = fmaxf(fmaxf(s_m[0], s_m[1]), fmaxf(s_m[2], s_m[3]));
            float ll = 0.f, o[VEC];
#pragma unroll
            for (int j = 0; j < VEC; ++j) o[j] = 0.f;
#pragma unroll
            for (int w = 0; w < 4; ++w) {
                const float c = (s_m[w] == -INFINITY) ? 0.f : exp2f(s_m[w] - mm);
                ll += s_l[w] * c;
#pragma unroll
                for (int j = 0; j < VEC; ++j) o[j] += s_acc[w][lane * VEC + j] * c;
            }
            half* op = P.out + (((size_t)b * P.q_len + i) * P.H + h) * hd + lane * VEC;
            const float inv = 1.f / ll;
#pragma unroll
            for (int j = 0; j < VEC; ++j) op[j] = __float2half_rn(o[j] * inv);
        }
        __syncthreads();
    }
}

}  // namespace exl2b

using namespace exl2b;

extern "C" int exl2b_paged_attn_decode(const uint16_t* q, const uint16_t* k_new, const uint16_t* v_new, uint16_t* k_cache,
                                       uint16_t* v_cache, const int32_t* cache_seqlens, const int32_t* block_table,
                                       uint16_t* out, int batch, int q_len, int num_heads, int num_kv_heads, int head_dim,
                                       int page_size, int pages_per_seq, float softmax_scale, exl2b_stream_t stream) {
    EXL2B_REQUIRE(q && k_new && v_new && k_cache && v_cache && cache_seqlens && block_table && out, "null argument");
    EXL2B_REQUIRE(head_dim == 64 || head_dim == 128, "head_dim %d not supported (64 or 128)", head_dim);
    EXL2B_REQUIRE(num_heads % num_kv_heads == 0, "bad GQA ratio");
    AttnParams P = {};
    P.q = (const half*)q; P.k_new = (const half*)k_new; P.v_new = (const half*)v_new;
    P.k_cache = (half*)k_cache; P.v_cache = (half*)v_cache;
    P.cache_seqlens = cache_seqlens; P.block_table = block_table; P.out = (half*)out;
    P.q_len = q_len; P.H = num_heads; P.KVH = num_kv_heads; P.hd = head_dim;
    P.page_size = page_size; P.pages_per_seq = pages_per_seq;
    P.scale_log2 = softmax_scale * 1.4426950408889634f;
    dim3 grid(num_heads, batch);
    if (head_dim == 128)
        EXL2B_CUDA(launch_pdl_f("attn", attn_decode_kernel<4>, grid, dim3(128), 0, (cudaStream_t)stream, P));
    else
        EXL2B_CUDA(launch_pdl_f("attn", attn_decode_kernel<2>, grid, dim3(128), 0, (cudaStream_t)stream, P));
    return 0;
}
