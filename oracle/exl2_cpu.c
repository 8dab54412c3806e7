/* C restatement of the reference's quantized GEMV for the CPU-baseline leg of bench.py.
 *
 * TEST INFRASTRUCTURE ONLY -- never linked into or called by the product (exllamav2_b200/).  bench.py's
 * `cpu_baseline` / `--impl reference` legs and tests/ are the only callers.  Parity: pinned through the numpy
 * oracle (oracle/exl2_oracle.py, itself pinned against the reference extension's outputs in tests/golden/);
 * tests/test_oracle_c.py checks this file against the numpy oracle.
 *
 * Semantics follow exllamav2_ext/cuda/q_gemm_kernel.cuh:140-565 (EXL2) and q_gemm_kernel_gptq.cuh:61-246 (GPTQ):
 *   y[n] = sum_g scale[g,n] * sum_{k' in g} a[perm[k']] * (q[k',n] - zero)
 * reading the CHECKPOINT layout directly (little-endian bit stream down K per column,
 * cuda/pack_tensor.cu:118-271), fp32 accumulation, all host threads (OpenMP over column chunks).
 */
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

static float half_to_float(uint16_t h) {
    uint32_t sign = (uint32_t)(h & 0x8000) << 16, exp = (h >> 10) & 0x1F, man = h & 0x3FF, f;
    if (exp == 0) {
        if (man == 0) f = sign;
        else { exp = 127 - 15 + 1; while (!(man & 0x400)) { man <<= 1; exp--; } man &= 0x3FF; f = sign | (exp << 23) | (man << 13); }
    } else if (exp == 31) f = sign | 0x7F800000u | (man << 13);
    else f = sign | ((exp + 127 - 15) << 23) | (man << 13);
    float r; memcpy(&r, &f, 4); return r;
}

int exl2_cpu_threads(void) {
#ifdef _OPENMP
    return omp_get_max_threads();
#else
    return 1;
#endif
}

/* EXL2.  q_weight uint32[R,N]; q_scale uint32[G,N/8]; q_scale_max fp16[G] (already * prescale/256);
 * q_groups int16[2G]; q_perm uint16[K] or NULL; a float[K] (ORIGINAL feature order); y float[N]. */
void exl2_cpu_gemv(const uint32_t* q_weight, const uint32_t* q_scale, const uint16_t* q_scale_max, const int16_t* q_groups,
                   const uint16_t* q_perm, int K, int N, int G, int R, const float* a, float* y) {
    float* ap = (float*)malloc(sizeof(float) * (size_t)K);
    for (int k = 0; k < K; ++k) ap[k] = a[q_perm ? q_perm[k] : k];
    const int CH = 256;
#pragma omp parallel for schedule(dynamic, 1)
    for (int n0 = 0; n0 < N; n0 += CH) {
        const int n1 = n0 + CH < N ? n0 + CH : N, w = n1 - n0;
        float acc[256], tot[256];
        for (int i = 0; i < w; ++i) tot[i] = 0.f;
        int k = 0;
        for (int g = 0; g < G; ++g) {
            const int bits = q_groups[2 * g], first = q_groups[2 * g + 1];
            const int next = (g + 1 < G) ? q_groups[2 * g + 3] : R;
            const int rows = (g + 1 < G) ? (next - first) * 32 / bits : K - k;
            const uint32_t mask = (1u << bits) - 1u;
            const float zp = (float)(1 << (bits - 1));
            for (int i = 0; i < w; ++i) acc[i] = 0.f;
            for (int r = 0; r < rows; ++r) {
                const int bitpos = r * bits, word = first + (bitpos >> 5), sh = bitpos & 31;
                const uint32_t* w0 = q_weight + (size_t)word * N + n0;
                const float av = ap[k + r];
                if (sh + bits <= 32) {
                    for (int i = 0; i < w; ++i) acc[i] += av * ((float)((w0[i] >> sh) & mask) - zp);
                } else {
                    const uint32_t* w1 = w0 + N;
                    for (int i = 0; i < w; ++i) acc[i] += av * ((float)(((w0[i] >> sh) | (w1[i] << (32 - sh))) & mask) - zp);
                }
            }
            const float smax = half_to_float(q_scale_max[g]);
            for (int i = 0; i < w; ++i) {
                const int n = n0 + i;
                const int qs = (int)((q_scale[(size_t)g * (N / 8) + (n >> 3)] >> ((n & 7) * 4)) & 15u) + 1;
                tot[i] += acc[i] * ((float)(qs * qs) * smax);
            }
            k += rows;
        }
        for (int i = 0; i < w; ++i) y[n0 + i] = tot[i];
    }
    free(ap);
}

/* GPTQ 4-bit without act-order (row k of the checkpoint = input feature k). */
void gptq_cpu_gemv(const uint32_t* qweight, const uint32_t* qzeros, const uint16_t* scales, int K, int N, int G,
                   const float* a, float* y) {
    int gs = 1;
    while (gs * G < K) gs *= 2;
    const int CH = 256;
#pragma omp parallel for schedule(dynamic, 1)
    for (int n0 = 0; n0 < N; n0 += CH) {
        const int n1 = n0 + CH < N ? n0 + CH : N, w = n1 - n0;
        float acc[256], tot[256], asum;
        for (int i = 0; i < w; ++i) tot[i] = 0.f;
        for (int g = 0; g < G; ++g) {
            for (int i = 0; i < w; ++i) acc[i] = 0.f;
            asum = 0.f;
            for (int k = g * gs; k < (g + 1) * gs && k < K; ++k) {
                const uint32_t* w0 = qweight + (size_t)(k >> 3) * N + n0;
                const int sh = (k & 7) * 4;
                const float av = a[k];
                asum += av;
                for (int i = 0; i < w; ++i) acc[i] += av * (float)((w0[i] >> sh) & 15u);
            }
            for (int i = 0; i < w; ++i) {
                const int n = n0 + i;
                const float z1 = (float)(((qzeros[(size_t)g * (N / 8) + (n >> 3)] >> ((n & 7) * 4)) & 15u) + 1);
                tot[i] += half_to_float(scales[(size_t)g * N + n]) * (acc[i] - z1 * asum);
            }
        }
        for (int i = 0; i < w; ++i) y[n0 + i] = tot[i];
    }
}
