/* exl2_b200.h -- C ABI of libexl2b200.so: the B200-native (sm_100a) quantized-linear hot path of ExLlamaV2.
 *
 * Drop-in boundary: the reference binds this path through the pybind11 module `exllamav2_ext`
 * (exllamav2/exllamav2_ext/ext_bindings.cpp:27-138).  Every entry point below names the reference binding it
 * replaces; exllamav2_b200/ext.py adapts torch.Tensor -> raw pointers and re-exports them under the reference's
 * own names, so exllamav2/{linear,attn,mlp,cache,rmsnorm}.py call it unchanged.
 *
 * Conventions
 *   - plain C: raw device pointers, ints, an opaque cudaStream_t passed as void*; no torch types.
 *   - every function returns 0 on success, non-zero on error; exl2b_last_error() returns the message
 *     (the reference raises C++ exceptions via TORCH_CHECK, cpp/util.h:34-39; the shim raises RuntimeError).
 *   - "absent tensor" (the reference's meta-device none_tensor, ext.py:296) is a NULL pointer.
 *   - all launches go to the given stream; nothing on the forward path synchronises the device.
 *   - fp16 tensors are `uint16_t*` (IEEE binary16 bit patterns) to keep the header free of cuda_fp16.h.
 */
#ifndef EXL2_B200_H
#define EXL2_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef void* exl2b_stream_t;   /* cudaStream_t */
typedef void* exl2b_qmatrix_t;  /* opaque; the reference's QMatrix* handle (cuda/q_matrix.cuh:11-83) */
typedef void* exl2b_qattn_t;    /* reference QAttn*  (cuda/q_attn.cuh:38-171) */
typedef void* exl2b_qmlp_t;     /* reference QMLP*   (cuda/q_mlp.cuh) */

const char* exl2b_last_error(void);
int exl2b_version(void);
/* Number of kernels this library has launched since load (bench.py reports it as gpu_launches). */
uint64_t exl2b_launch_count(void);

/* ---- QMatrix ------------------------------------------------------------------------------------------------
 * replaces make_q_matrix (ext_qmatrix.cpp:21-111) / QMatrix::QMatrix (cuda/q_matrix.cu:49-196).
 * Exactly one of {q_scale,...} (EXL2) or {gptq_qzeros,...} (GPTQ) is non-NULL.
 * Like the reference (shuffle_kernel, q_matrix.cu:187-195) creation REWRITES q_weight in place into a private
 * layout (when width % 64 == 0; otherwise the handle owns a padded copy); q_scale/q_scale_max/qzeros/scales/
 * q_perm/bias stay owned by the caller and must outlive the handle (linear.py:145 keeps them alive).
 * q_scale_max must already carry the prescale/256 factor (ext.py:336).  q_groups may be a device or host pointer
 * (int16[2*groups]); gptq_g_idx is a HOST pointer (the reference passes w["g_idx"].cpu(), ext.py:385) or NULL;
 * for act-order GPTQ q_perm/q_invperm are caller-allocated device buffers that are FILLED here (ext.py:373-374,
 * q_matrix.cu:646-647).  temp_dq/max_dq_rows of the reference are not needed (no reconstruct+cuBLAS detour). */
typedef struct exl2b_qmatrix_desc {
    int device;
    int height;            /* K = in_features  */
    int width;             /* N = out_features */
    int groups;
    uint32_t* q_weight;    /* EXL2: int32[R, N];  GPTQ: qweight int32[K/8, N] */
    uint16_t* q_perm;      /* int16[K] or NULL (stored row k' <- input feature q_perm[k']) */
    uint16_t* q_invperm;   /* int16[K] or NULL */
    const uint32_t* q_scale;      /* EXL2 int32[G, N/8] */
    const uint16_t* q_scale_max;  /* EXL2 fp16[G], pre-multiplied by prescale/256 */
    const uint16_t* q_groups;     /* EXL2 int16[2G] (bits, first packed row) */
    int q_weight_rows;            /* EXL2: R (rows of q_weight) */
    const uint32_t* gptq_qzeros;  /* GPTQ int32[G, N/8] */
    const uint16_t* gptq_scales;  /* GPTQ fp16[G, N] */
    const int32_t* gptq_g_idx;    /* GPTQ int32[K] on the HOST, or NULL */
    const uint16_t* bias;         /* fp16[N] or NULL */
} exl2b_qmatrix_desc;

int exl2b_qmatrix_create(const exl2b_qmatrix_desc* desc, exl2b_stream_t stream, exl2b_qmatrix_t* out);
int exl2b_qmatrix_destroy(exl2b_qmatrix_t h);                     /* free_q_matrix, ext_qmatrix.cpp:187-194 */
int exl2b_qmatrix_info(exl2b_qmatrix_t h, int* height, int* width, int* groups, int* is_gptq, uint64_t* packed_bytes);

/* reconstruct (ext_qmatrix.cpp:196-210, QMatrix::reconstruct q_matrix.cu:499-553):
 * out fp16[K, N] row-major in ORIGINAL row order, out[perm[k'], n] = half(q - zero) * half(scale), bit-exact. */
int exl2b_reconstruct(exl2b_qmatrix_t h, uint16_t* out, exl2b_stream_t stream);

/* gemm_half_q_half (ext_qmatrix.cpp:213-247 -> gemm_half_q_half_cuda, cuda/q_gemm.cu:201-313):
 * c[m, n] = (clear ? 0 : c[m, n]) + bias[n] + sum_k a[m, k] * W[k, n];  a fp16[M, K] (lda = row stride in
 * elements), c fp16[M, N] (ldc).  All M are served by the fused dequant kernels (no temp_dq, no cuBLAS);
 * force_cuda is accepted for signature parity and ignored. */
int exl2b_gemm_half_q_half(exl2b_qmatrix_t h, const uint16_t* a, int lda, uint16_t* c, int ldc, int m, int clear,
                           int force_cuda, exl2b_stream_t stream);

/* rms_norm (ext_norm.cpp:23-62) + gemm_half_q_half on ONE row as a single launch -- the final norm + lm_head of a decode
 * step (exllamav2/model.py:1036-1044 runs them as two kernels): c[n] = (clear ? 0 : c[n]) + bias[n] +
 * sum_k half(x[k] * w[k] * rsqrt(mean(x^2) + eps)) * W[k, n];  x fp16[K], c fp16[N]. */
int exl2b_gemm_half_q_half_norm(exl2b_qmatrix_t h, const uint16_t* x, const uint16_t* norm_w, float norm_eps, uint16_t* c,
                                int clear, exl2b_stream_t stream);

/* make_group_map (ext_qmatrix.cpp:341-361): host-only helper, out int16[2*K]; returns rows written / 2 in *k. */
int exl2b_make_group_map(const int16_t* q_groups, int num_groups, int num_qrows, int16_t* out, int out_capacity, int* k);

/* ---- RMSNorm / RoPE / activation ----------------------------------------------------------------------------
 * rms_norm / rms_norm_ (ext_norm.cpp:23-103 -> rms_norm_cuda, cuda/rms_norm.cu:177-229): y = x*w*rsqrt(mean(x^2)+eps),
 * fp16 in/out (y may alias x). */
int exl2b_rms_norm(const uint16_t* x, const uint16_t* w, uint16_t* y, float eps, int rows, int dim, exl2b_stream_t stream);

/* rope_ (ext_rope.cpp:20-62 -> rope_cuda, cuda/rope.cu:220-273): in-place rotary on x fp16[batch, rows_per_batch,
 * head_dim] where row = token*num_heads + head; position = past_len (+ past_lens[b], or past_lens[b] alone when
 * past_len == -1) + row / num_heads.  sin/cos fp16[max_pos, sincos_size]; neox != 0 -> half-split pairs. */
int exl2b_rope(uint16_t* x, const uint16_t* sin, const uint16_t* cos, int batch, int rows_per_batch, int head_dim,
               int num_heads, int past_len, const int32_t* past_lens, int neox, int sincos_size, exl2b_stream_t stream);

/* act_mul (cuda/q_mlp_activation.cuh:54-196): x = silu(x) * y (gelu when act_gelu), fp16, in place on x. */
int exl2b_act_mul(uint16_t* x, const uint16_t* y, int rows, int width, int act_gelu, exl2b_stream_t stream);

/* ---- Q4 K/V cache -------------------------------------------------------------------------------------------
 * fp16_to_q_kv / q_to_fp16_kv (ext_cache.cpp:80-274 -> cuda/cache.cu:143-497, cuda/cache_q.cuh), wbits = 4.
 * Non-paged (page_size == 0): tensors are [batch, seq, heads*head_dim]; `dim` = heads*head_dim, `seq_stride` =
 * elements per batch row of the fp16 tensor; tokens [offset, offset+width) of every batch row are converted.
 * Paged (page_size > 0): block_table int32[batch, pages_per_seq], cache_seqlens int32[batch]; pack converts
 * tokens [seqlen, seqlen+q_len) (q_len passed in `width`), unpack converts [0, seqlen) of every sequence.
 * v_* may be NULL (single tensor). */
int exl2b_fp16_to_q_kv(const uint16_t* k_in, uint8_t* k_out, uint16_t* k_scales,
                       const uint16_t* v_in, uint8_t* v_out, uint16_t* v_scales,
                       int batch, int dim, int seq_stride, int offset, int width,
                       int page_size, const int32_t* cache_seqlens, const int32_t* block_table, int pages_per_seq,
                       int wbits, exl2b_stream_t stream);
int exl2b_q_to_fp16_kv(const uint8_t* k_in, const uint16_t* k_scales, uint16_t* k_out,
                       const uint8_t* v_in, const uint16_t* v_scales, uint16_t* v_out,
                       int batch, int dim, int seq_stride, int offset, int width,
                       int page_size, const int32_t* cache_seqlens, const int32_t* block_table, int pages_per_seq,
                       int wbits, exl2b_stream_t stream);

/* ---- fused attention / MLP blocks ---------------------------------------------------------------------------
 * make_q_attn / q_attn_forward_1 / q_attn_forward_2 (ext_qattn.cpp:24-191 -> cuda/q_attn.cu:153-345).
 * forward_1: RMSNorm(x) -> Q,K,V projections -> RoPE on Q and K; forward_2: x (+)= attn_out @ o_proj.
 * RMSNorm is folded into the projection kernel's prologue and Q/K/V run as ONE launch. */
typedef struct exl2b_qattn_desc {
    const uint16_t* layernorm;   /* fp16[hidden] RMSNorm weight or NULL */
    float norm_epsilon;
    exl2b_qmatrix_t q_proj, k_proj, v_proj, o_proj;
    int hidden_size, num_heads, num_kv_heads, head_dim;
    int has_residual;
    int rope_style;              /* 0 none, 1 gptj, 2 neox  (reference ROPE_STYLE_*, cuda/rope.cuh) */
    int sincos_size;
} exl2b_qattn_desc;
int exl2b_qattn_create(const exl2b_qattn_desc* desc, exl2b_qattn_t* out);
int exl2b_qattn_destroy(exl2b_qattn_t h);
int exl2b_qattn_forward_1(exl2b_qattn_t h, const uint16_t* x, int batch, int q_len, int past_len, const int32_t* past_lens,
                          uint16_t* q, uint16_t* k, uint16_t* v, const uint16_t* sin, const uint16_t* cos, exl2b_stream_t stream);
int exl2b_qattn_forward_2(exl2b_qattn_t h, uint16_t* x, const uint16_t* attn_out, int batch, int q_len, exl2b_stream_t stream);

/* make_q_mlp / q_mlp_forward_ (ext_qmlp.cpp:22-118 -> QMLP::forward_, cuda/q_mlp.cu:78-236):
 * x (+)= down( act(gate(norm(x))) * up(norm(x)) );  temp_a/temp_b fp16[rows, intermediate] scratch. */
typedef struct exl2b_qmlp_desc {
    const uint16_t* layernorm;
    float norm_epsilon;
    exl2b_qmatrix_t gate, up, down;
    int hidden_size, intermediate_size;
    int act_gelu;
    int has_residual;
} exl2b_qmlp_desc;
int exl2b_qmlp_create(const exl2b_qmlp_desc* desc, exl2b_qmlp_t* out);
int exl2b_qmlp_destroy(exl2b_qmlp_t h);
int exl2b_qmlp_forward(exl2b_qmlp_t h, uint16_t* x, int rows, uint16_t* temp_a, uint16_t* temp_b, exl2b_stream_t stream);
/* Column-sharded (tensor-parallel) use, exllamav2/tensor_p.py + ext_qattn.cpp:261-732 / ext_qmlp.cpp:326-473 in the
 * reference: a rank creates the attention block with ITS heads (num_heads / num_kv_heads = local counts) and o_proj = NULL,
 * the MLP block with ITS intermediate columns and down = NULL, runs part 1 / the gate|up half locally, all-gathers the
 * sharded activation, and applies its column shard of o_proj / down with exl2b_gemm_half_q_half(clear = 0) into its slice of
 * the residual stream (ldc = hidden size). */
int exl2b_qmlp_forward_gateup(exl2b_qmlp_t h, const uint16_t* x, int rows, uint16_t* temp_a, exl2b_stream_t stream);

/* ---- host-buffer entry point (bench.py "e2e"): a fp16[M,K] and c fp16[M,N] are HOST pointers (pinned or
 * pageable); the call copies a to the device, runs gemm_half_q_half, copies c back and waits. */
int exl2b_gemm_half_q_half_host(exl2b_qmatrix_t h, const uint16_t* a_host, uint16_t* c_host, int m, exl2b_stream_t stream);

/* Tuning / diagnostics hook (no reference counterpart): CTAs per SM of the GEMV grid (<= 0 keeps the current value)
 * and an optional device buffer of 64 x 8 uint64 (one slot per launch, round robin): [0..5] globaltimer phase stamps of
 * CTA `cta`, [6] earliest CTA start (pre-fill with ~0), [7] latest CTA end (NULL disables). */
int exl2b_debug_set(int ctas_per_sm, unsigned long long* stamps, int cta);

/* Host-only diagnostics (no GPU needed): the 32-column-block -> CTA partition the batch-1 GEMV uses for blocks of the given
 * byte sizes; out[0..*used] are the block boundaries of the CTAs. */
int exl2b_debug_partition(const uint32_t* block_bytes, int num_blocks, int ctas, uint16_t* out, int* used);
/* diagnostics: per-CTA records (start, wait over, end, SM id) of the batch-1 GEMV launches, [64][160][4] uint64, or NULL */
int exl2b_debug_set_records(unsigned long long* records);
/* host-only: the per-warp stage lists of the batch-1 GEMV for one matrix structure (tests/test_i8_emulation.py) */
int exl2b_debug_plan(int N, int KS, int is_gptq, uint32_t blk_stream_bytes, const int* regions, int num_regions, int ctas, int warps,
                     int slot_bytes, uint32_t* desc, int cap_desc, uint32_t* first, int* ctas_used, int* n_desc, int* lcap,
                     uint32_t* red);

/* Stand-in for flash_attn_with_kvcache (third-party in the reference, attn.py:602-613): appends the q_len new K/V rows
 * to the paged fp16 cache at [seqlen, seqlen+q_len) and attends causally.  q [batch,q_len,H,hd], k/v_new
 * [batch,q_len,KVH,hd], caches fp16 [pages,page_size,KVH,hd], out [batch,q_len,H,hd]; head_dim 64 or 128. */
int exl2b_paged_attn_decode(const uint16_t* q, const uint16_t* k_new, const uint16_t* v_new, uint16_t* k_cache,
                            uint16_t* v_cache, const int32_t* cache_seqlens, const int32_t* block_table, uint16_t* out,
                            int batch, int q_len, int num_heads, int num_kv_heads, int head_dim, int page_size,
                            int pages_per_seq, float softmax_scale, exl2b_stream_t stream);

/* Decode attention straight over the Q4 cache (replaces the reference's per-layer sequence q_to_fp16_kv ->
 * flash_attn_with_kvcache -> fp16_to_q_kv, attn.py:560-613 + cache.py:472-556): quantises the q_len new K/V rows with
 * the fp16_to_q_kv arithmetic, appends them to the paged Q4 cache at [seqlen, seqlen+q_len) and attends causally over
 * the stored 4-bit values.  k/v_cache uint8 [pages,page_size,KVH,hd/2], k/v_scales fp16 [pages,page_size,KVH,hd/32];
 * 1 <= q_len <= 8; head_dim 64 or 128. */
int exl2b_paged_attn_decode_q4(const uint16_t* q, const uint16_t* k_new, const uint16_t* v_new, uint8_t* k_cache,
                               uint16_t* k_scales, uint8_t* v_cache, uint16_t* v_scales, const int32_t* cache_seqlens,
                               const int32_t* block_table, uint16_t* out, int batch, int q_len, int num_heads,
                               int num_kv_heads, int head_dim, int page_size, int pages_per_seq, float softmax_scale,
                               exl2b_qmatrix_t out_consumer, exl2b_stream_t stream);

/* Same, with RoPE fused: when rope_style != 0 (1 gptj, 2 neox) and the tables are given, q and k_new are the UN-rotated
 * projection outputs and are rotated as they are read (rope_cuda arithmetic, cuda/rope.cu:52-67,111-122; position of row i =
 * cache_seqlens[b] + i), so q_attn_forward_1 can skip its rope launches (pass sin = cos = NULL there).  Needs
 * sincos_size == head_dim.  With `out_consumer` and a single row the attention output is also left, as a plain fp16 row in
 * the consumer's stored-row order, where the batch-1 GEMV of o_proj reads it. */
int exl2b_paged_attn_decode_q4_ex(const uint16_t* q, const uint16_t* k_new, const uint16_t* v_new, uint8_t* k_cache,
                                  uint16_t* k_scales, uint8_t* v_cache, uint16_t* v_scales, const int32_t* cache_seqlens,
                                  const int32_t* block_table, uint16_t* out, int batch, int q_len, int num_heads,
                                  int num_kv_heads, int head_dim, int page_size, int pages_per_seq, float softmax_scale,
                                  exl2b_qmatrix_t out_consumer, const uint16_t* rope_sin, const uint16_t* rope_cos,
                                  int rope_style, int sincos_size, exl2b_stream_t stream);
/* Sticky status of the fused attention kernels on `device` (synchronises): bit 0 = a sequence would have run past its page
 * table (cache_seqlens[b] + q_len > pages_per_seq * page_size); that call appended nothing and wrote no output. */
int exl2b_paged_attn_status(int device, int* status);
int exl2b_paged_attn_clear_status(int device);

/* ---- chained launches (no reference counterpart; the reference runs norm / projection / rope / activation as separate
 * kernels, q_attn.cu:153-345, q_mlp.cu:78-236).  A producer's epilogue can write its output straight into the
 * activation buffer of the matrices that consume it -- permuted through their q_invperm, in the tensor-core operand
 * layout, pre-multiplied by the RMSNorm weight they apply -- together with per-strip sums of squares; the consumer
 * launch (`input_prepared` = 1) then starts without a prep kernel and applies 1/rms to its fp32 result.
 * Valid for rows <= 8 (decode) and the default (tcgen05) matrix layout; the calls fail otherwise.
 * `out_consumer` of exl2b_paged_attn_decode_q4 is the same mechanism for the attention output (o_proj). */
typedef struct {
    exl2b_qmatrix_t consumers[3];   /* matrices whose INPUT is this launch's output (e.g. the next block's q, k, v) */
    int num_consumers;
    const uint16_t* norm_weight;    /* RMSNorm weight the consumers apply to that input, or NULL */
} exl2b_chain_t;
int exl2b_qattn_forward_1_ex(exl2b_qattn_t h, const uint16_t* x, int batch, int q_len, int past_len, const int32_t* past_lens,
                             uint16_t* q, uint16_t* k, uint16_t* v, const uint16_t* sin, const uint16_t* cos,
                             int input_prepared, exl2b_stream_t stream);
int exl2b_qattn_forward_2_ex(exl2b_qattn_t h, uint16_t* x, const uint16_t* attn_out, int batch, int q_len, int input_prepared,
                             const exl2b_chain_t* next, exl2b_stream_t stream);
int exl2b_qmlp_forward_ex(exl2b_qmlp_t h, uint16_t* x, int rows, uint16_t* temp_a, uint16_t* temp_b, int input_prepared,
                          const exl2b_chain_t* next, exl2b_stream_t stream);
/* gemm_half_q_half on an input prepared by a chained producer (lm_head after the last MLP; has_norm: apply 1/rms) */
int exl2b_gemm_half_q_half_prepared(exl2b_qmatrix_t h, uint16_t* c, int ldc, int m, int clear, int has_norm, float norm_eps,
                                    exl2b_stream_t stream);
int exl2b_qmatrix_chain_target(exl2b_qmatrix_t h, uint16_t** xp, const uint16_t** invperm);

#ifdef __cplusplus
}
#endif
#endif /* EXL2_B200_H */
