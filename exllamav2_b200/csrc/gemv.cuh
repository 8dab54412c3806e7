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
    int tc_act_bytes;      // tcgen05 kernel: capacity of the staged activations (16 B per k)
};

// Launch one or more passes (8 tokens each) of the GEMV over `nm` matrices that share K and the input layout.
// All matrices must live on `device`.  M may exceed 8 (extra passes re-read the weights, like the reference's
// grid.y = ceil(M/4) does, cuda/q_gemm.cu:97).
int gemv_launch(int device, cudaStream_t stream, GemvMat* mats, int nm, int M, const half* norm_w, float norm_eps,
                int epilogue);

}  // namespace exl2b
