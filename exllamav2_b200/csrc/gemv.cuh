// Internal interface of the streaming dequant-GEMV (gemv.cu), shared with the fused attention / MLP blocks.
#pragma once
#include "qmatrix.cuh"

namespace exl2b {

constexpr int GEMV_MAX_MATS = 3;
constexpr int GEMV_MTOK = 8;          // tokens per pass (the N=8 dimension of mma.m16n8k16)

enum GemvEpilogue : int {
    EPI_STORE = 0,      // c = (clear ? 0 : c) + bias + acc
    EPI_SILU_MUL = 1,   // mats = {gate, up}: c0 = silu(gate) * up   (written to mat[0].c only)
    EPI_GELU_MUL = 2,
};

struct GemvMat {
    QMatView w;
    const half* xp;  // tcgen05 path: activations prepared by tc_prep_kernel (normalised, permuted, UMMA core-matrix layout)
    const half* x;   // input activations fp16 [M][ldx], ORIGINAL feature order (the kernel gathers through w.perm)
    int ldx;
    half* c;         // output fp16 [M][ldc]
    int ldc;
    int clear;       // 1: overwrite, 0: accumulate into c (residual add, cuda/q_attn.cu:333)
    int unit_begin;  // filled by the launcher
    int strip_begin; // filled by the launcher
};

// ---- optional fusions around a launch (tcgen05 path) --------------------------------------------------------------------
// RoPE applied in the epilogue of the matrices selected by `mask` (cuda/rope.cu:10-123 arithmetic; needs 128 % head_dim == 0
// so that a rotation partner lives in the same 128-column strip)
struct RopeFuse {
    const half* sin;
    const half* cos;
    const int32_t* past_lens;
    int past_len, q_len, head_dim, sincos_size, neox;
    unsigned mask;          // bit i: rotate mat[i]'s output
};
// A consumer of this launch's output: its activation buffer is filled directly from the epilogue, already permuted
// into the consumer's row order, in the UMMA core-matrix layout, times the consumer's RMSNorm weight -- the consumer
// launch then needs no prep kernel.  The RMSNorm's 1/rms is deferred: the producer leaves per-strip sums of squares
// in `sumsq`, the consumer multiplies its fp32 result by rsqrt(sum / K + eps).
struct ScatterTarget {
    half* xp;
    const uint16_t* invperm;    // consumer row of feature n (NULL: identity)
    const half* scale;          // RMSNorm weight (NULL: none)
};
struct GemvExtras {
    RopeFuse rope;                         // mask == 0: none
    ScatterTarget scat[GEMV_MAX_MATS];
    int num_scat;
    float* sumsq_out;                      // [strips][8] or NULL
    int prepared;                          // 1: mats[i].xp already hold the input (no prep launch)
    const float* sumsq_in;                 // with prepared: per-strip sums of squares of the input rows, or NULL (no norm)
    int sumsq_in_strips;
    float sumsq_eps;
};

struct GemvParams {
    GemvMat mat[GEMV_MAX_MATS];
    int num_mats;
    int M;                 // tokens this pass, 1..8
    int KS;                // slabs per strip (K/32), common to all matrices of the launch
    int total_units;       // sum over matrices of strips*KS
    const half* norm_w;    // fused RMSNorm weight (NULL: none): a = half(x * w * rsqrt(mean(x^2)+eps))
    float norm_eps;
    int epilogue;
    float* ws;             // split-K partial sums
    unsigned int* counters;
    int maxc;              // max contributors per strip (workspace stride)
    int act_stride;        // bytes between token rows of the staged activations in shared memory
    int act_rows;          // rows staged per segment (capacity)
    unsigned long long* dbg;   // optional phase timestamps (globaltimer) of CTA dbg_cta, NULL in production
    int dbg_cta;
    int tc_stage_bytes;    // tcgen05 kernel: bytes of one weight stage (largest group of one 32-column block)
    int tc_act_off;        // tcgen05 kernel: shared-memory offset of the staged activations
    int tc_act_bytes;      // tcgen05 kernel: bytes of the two activation rings
    int tc_stages;         // tcgen05 kernel: pipeline stages (groups in flight per warp), 2..4
    int row0;              // first token row of this pass (RoPE position bookkeeping)
    GemvExtras ex;
};

// Launch one or more passes (8 tokens each) of the GEMV over `nm` matrices that share K and the input layout.
// All matrices must live on `device`.  M may exceed 8 (extra passes re-read the weights, like the reference's
// grid.y = ceil(M/4) does, cuda/q_gemm.cu:97).
int gemv_launch(int device, cudaStream_t stream, GemvMat* mats, int nm, int M, const half* norm_w, float norm_eps,
                int epilogue, const GemvExtras* ex = nullptr);
// Many-row path (gemm_big.cu): reconstruct a column window + cuBLAS fp16 GEMM with fp32 accumulation.  Rows above
// GEMM_BIG_MIN_ROWS take it (below, the packed-row kernels re-read the weights at most twice).
constexpr int GEMM_BIG_MIN_ROWS = 16;
bool gemm_big_available();
int gemm_big_launch(const QMatrix* q, const half* a, int lda, half* c, int ldc, int M, int clear, cudaStream_t stream);

// can `ex` be honoured for these matrices / this row count?  (tcgen05 layout, one pass)
bool gemv_supports_extras(const GemvMat* mats, int nm, int M);

}  // namespace exl2b
