// Many-row (prefill) path of gemm_half_q_half: dequantise a column window of the matrix into a stream-ordered fp16 temp
// (reconstruct_kernel, bit-exact with the reference's reconstruct) and hand the dense contraction to cuBLAS on the tensor
// cores -- the regime switch the reference makes in gemm_half_q_half_cuda for more than MAX_Q_GEMM_ROWS rows
// (exllamav2_ext/cuda/q_gemm.cu:233-266: reconstruct into temp_dq, cublasHgemm), with fp32 accumulation instead of fp16.
// Above ~16 rows every weight is re-used often enough that the temp traffic (2 B written + 2 B read per weight per window)
// is noise next to the packed-row passes of the decode kernels (one full weight read per 8 rows).
//
// cuBLAS is a plain library GEMM here (dense fp16 x fp16 -> fp16, no fusion needed) and is bound at RUN time with dlopen:
// libexl2b200.so keeps no link-time dependency beyond the CUDA runtime, and the decode path never touches it.
#include <cublas_v2.h>
#include <dlfcn.h>

#include <algorithm>
#include <mutex>

#include "qmatrix.cuh"

namespace exl2b {

int reconstruct_window(const QMatrix* m, half* out, int ld, int strip0, int nstrips, cudaStream_t stream);

namespace {

struct Cublas {
    void* lib = nullptr;
    cublasStatus_t (*create)(cublasHandle_t*) = nullptr;
    cublasStatus_t (*set_stream)(cublasHandle_t, cudaStream_t) = nullptr;
    cublasStatus_t (*set_workspace)(cublasHandle_t, void*, size_t) = nullptr;
    cublasStatus_t (*gemm_ex)(cublasHandle_t, cublasOperation_t, cublasOperation_t, int, int, int, const void*, const void*, cudaDataType,
                              int, const void*, cudaDataType, int, const void*, void*, cudaDataType, int, cublasComputeType_t,
                              cublasGemmAlgo_t) = nullptr;
    cublasHandle_t handle[64] = {nullptr};
    void* workspace[64] = {nullptr};
    bool tried = false, ok = false;
};
Cublas g_cb;
std::mutex g_cb_mutex;

bool cublas_load() {
    if (g_cb.tried) return g_cb.ok;
    g_cb.tried = true;
    for (const char* name : {"libcublas.so.12", "/usr/local/cuda/lib64/libcublas.so.12", "libcublas.so"}) {
        g_cb.lib = dlopen(name, RTLD_NOW | RTLD_GLOBAL);
        if (g_cb.lib) break;
    }
    if (!g_cb.lib) return false;
    g_cb.create = (decltype(g_cb.create))dlsym(g_cb.lib, "cublasCreate_v2");
    g_cb.set_stream = (decltype(g_cb.set_stream))dlsym(g_cb.lib, "cublasSetStream_v2");
    g_cb.set_workspace = (decltype(g_cb.set_workspace))dlsym(g_cb.lib, "cublasSetWorkspace_v2");
    g_cb.gemm_ex = (decltype(g_cb.gemm_ex))dlsym(g_cb.lib, "cublasGemmEx");
    g_cb.ok = g_cb.create && g_cb.set_stream && g_cb.set_workspace && g_cb.gemm_ex;
    return g_cb.ok;
}

constexpr size_t BIG_WORKSPACE = (size_t)32 << 20;
constexpr size_t BIG_TEMP_BYTES = (size_t)64 << 20;       // dequantised window: K x columns x 2 B

__global__ void bias_rows_kernel(half* __restrict__ c, int ldc, const half* __restrict__ bias, int n0, int ncols, int rows, int set) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= rows * ncols) return;
    const int r = i / ncols, n = i - r * ncols;
    half* p = c + (size_t)r * ldc + n0 + n;
    *p = set ? bias[n0 + n] : __hadd(*p, bias[n0 + n]);
}

}  // namespace

bool gemm_big_available() {
    std::lock_guard<std::mutex> lk(g_cb_mutex);
    return cublas_load();
}

// c[M, N] = (clear ? 0 : c) + bias + a[M, K] @ W[K, N]
int gemm_big_launch(const QMatrix* q, const half* a, int lda, half* c, int ldc, int M, int clear, cudaStream_t stream) {
    const QMatView& v = q->v;
    cublasHandle_t h = nullptr;
    // one cuBLAS handle per device: host-side use is serialised (the work itself is asynchronous on `stream`)
    std::lock_guard<std::mutex> lk(g_cb_mutex);
    {
        EXL2B_REQUIRE(cublas_load(), "cuBLAS is not loadable (libcublas.so.12): the many-row path is unavailable");
        const int d = q->device;
        EXL2B_REQUIRE(d >= 0 && d < 64, "bad device");
        if (!g_cb.handle[d]) {
            EXL2B_REQUIRE(g_cb.create(&g_cb.handle[d]) == CUBLAS_STATUS_SUCCESS, "cublasCreate failed");
            EXL2B_CUDA(cudaMalloc(&g_cb.workspace[d], BIG_WORKSPACE));
        }
        h = g_cb.handle[d];
        EXL2B_REQUIRE(g_cb.set_stream(h, stream) == CUBLAS_STATUS_SUCCESS, "cublasSetStream failed");
        // (set after the stream: cuBLAS resets the workspace on cublasSetStream) a fixed workspace keeps the call capturable
        g_cb.set_workspace(h, g_cb.workspace[d], BIG_WORKSPACE);
    }
    const int SN = strip_n(v.layout);
    const int max_strips = std::max<int>(1, (int)(BIG_TEMP_BYTES / ((size_t)v.K * SN * sizeof(half))));
    half* temp = nullptr;
    const int win_strips = std::min(max_strips, v.strips);
    EXL2B_CUDA(cudaMallocAsync(&temp, (size_t)v.K * win_strips * SN * sizeof(half), stream));
    int rc = 0;
    for (int s0 = 0; s0 < v.strips && rc == 0; s0 += win_strips) {
        const int ns = std::min(win_strips, v.strips - s0);
        const int col0 = s0 * SN, ncols = std::min(ns * SN, v.N - col0);
        rc = reconstruct_window(q, temp, ncols, s0, ns, stream);
        if (rc) break;
        float beta = clear ? 0.f : 1.f;
        if (v.bias && clear) {      // bias first, then accumulate the product onto it
            bias_rows_kernel<<<(M * ncols + 255) / 256, 256, 0, stream>>>(c, ldc, v.bias, col0, ncols, M, 1);
            g_launch_count++;
            beta = 1.f;
        }
        const float alpha = 1.f;
        // row-major C[M, ncols] = A[M, K] W[K, ncols]  ==  column-major C^T[ncols, M] = W^T[ncols, K] A^T[K, M]
        const cublasStatus_t st = g_cb.gemm_ex(h, CUBLAS_OP_N, CUBLAS_OP_N, ncols, M, v.K, &alpha, temp, CUDA_R_16F, ncols, a, CUDA_R_16F, lda,
                                               &beta, c + col0, CUDA_R_16F, ldc, CUBLAS_COMPUTE_32F, CUBLAS_GEMM_DEFAULT);
        if (st != CUBLAS_STATUS_SUCCESS) {
            set_error("cublasGemmEx failed with status %d", (int)st);
            rc = -1;
            break;
        }
        if (v.bias && !clear) {
            bias_rows_kernel<<<(M * ncols + 255) / 256, 256, 0, stream>>>(c, ldc, v.bias, col0, ncols, M, 0);
            g_launch_count++;
        }
    }
    cudaFreeAsync(temp, stream);
    return rc;
}

}  // namespace exl2b
