// Dispatch of gemm_half_q_half over the row count, and the split-K workspace of the tcgen05 kernel.
//   1 row        -> gemv_i8.cu   (HBM-bound integer GEMV, the decode path)
//   2 .. 16 rows -> gemm_tc.cu   (packed weights as the tcgen05 A operand from tensor memory, 8 rows per pass)
//   more         -> gemm_big.cu  (reconstruct window + dense tensor-core GEMM: the reference's regime above MAX_Q_GEMM_ROWS)
// Replaces gemm_half_q_half_cuda (exllamav2_ext/cuda/q_gemm.cu:201-313).
#include <algorithm>
#include <map>
#include <mutex>

#include "gemv.cuh"
#include "gemv_i8.cuh"

namespace exl2b {

int g_ctas_per_sm = 2;
int g_tc_ctas_per_sm = [] { const char* e = getenv("EXL2B_TC_CTAS"); return e ? atoi(e) : 2; }();
unsigned long long* g_dbg = nullptr;
int g_dbg_cta = 0;
int g_dbg_slot = 0;
unsigned long long* g_dbg_rec = nullptr;      // optional per-CTA records of the batch-1 GEMV: [64 launches][160 CTAs][4]

// Split-K workspace, arrival counters and the activation-operand scratch of the tcgen05 kernel, one set per (device, stream):
// launches on different streams (or host threads driving different streams) never share scratch.  Created on first use --
// never inside a stream capture: call once eagerly first, as model.capture() does.
struct TcWorkspace {
    float* ws = nullptr;
    unsigned int* counters = nullptr;
    half* xp = nullptr;
    size_t ws_bytes = 0, xp_bytes = 0;
    int n_counters = 0;
};
static std::map<std::pair<int, cudaStream_t>, TcWorkspace> g_tc_ws;
static std::mutex g_ws_mutex;

int gemv_workspace(int device, cudaStream_t stream, float** ws, unsigned int** counters, size_t* ws_bytes, int* n_counters, half** xp,
                   size_t xp_bytes) {
    std::lock_guard<std::mutex> lock(g_ws_mutex);
    TcWorkspace& d = g_tc_ws[{device, stream}];
    if (!d.ws) {
        d.ws_bytes = (size_t)32 << 20;
        d.n_counters = 1 << 16;
        EXL2B_CUDA(cudaMalloc(&d.ws, d.ws_bytes));
        EXL2B_CUDA(cudaMalloc(&d.counters, d.n_counters * sizeof(unsigned int)));
        EXL2B_CUDA(cudaMemset(d.counters, 0, d.n_counters * sizeof(unsigned int)));
        EXL2B_CUDA(cudaDeviceSynchronize());
    }
    if (d.xp_bytes < xp_bytes) {
        if (d.xp) EXL2B_CUDA(cudaFree(d.xp));
        EXL2B_CUDA(cudaMalloc(&d.xp, xp_bytes));
        d.xp_bytes = xp_bytes;
    }
    *ws = d.ws;
    *counters = d.counters;
    *ws_bytes = d.ws_bytes;
    *n_counters = d.n_counters;
    *xp = d.xp;
    return 0;
}

int gemm_tc_launch(int device, cudaStream_t stream, GemvMat* mats, int nm, int M, const half* norm_w, float norm_eps, int epilogue,
                   const GemvExtras* ex);

bool gemm_tc_supported(const QMatView& v);
bool gemv_supports_extras(const GemvMat* mats, int nm, int M) {
    for (int i = 0; i < nm; ++i)
        if (mats[i].w.layout != LAYOUT_TC || !gemm_tc_supported(mats[i].w)) return false;
    return M >= 1 && M <= GEMV_MTOK;
}

int gemv_launch(int device, cudaStream_t stream, GemvMat* mats, int nm, int M, const half* norm_w, float norm_eps,
                int epilogue, const GemvExtras* ex) {
    EXL2B_REQUIRE(nm >= 1 && nm <= GEMV_MAX_MATS, "bad matrix count %d", nm);
    EXL2B_REQUIRE(device >= 0 && device < 64, "bad device %d", device);
    if (M <= 0) return 0;
    for (int i = 0; i < nm; ++i) EXL2B_REQUIRE(mats[i].w.layout == LAYOUT_TC, "matrix is not in the default (tcgen05) layout");
    return gemm_tc_launch(device, stream, mats, nm, M, norm_w, norm_eps, epilogue, ex);
}

// The tcgen05 kernel stages one quantisation group of a 32-column block per ring slot (<= 4 KB): groups of 256+ rows above 4
// bits, or ungrouped GPTQ, do not fit.  Such matrices take the dense path for every row count above one.
bool gemm_tc_supported(const QMatView& v) {
    for (int r = 0; r < v.num_regions; ++r)
        if ((1 << v.reg[r].spg_log2) * block_bytes(v.reg[r].bits) > 4096) return false;
    return true;
}

}  // namespace exl2b

using namespace exl2b;

extern "C" int exl2b_gemm_half_q_half(exl2b_qmatrix_t h, const uint16_t* a, int lda, uint16_t* c, int ldc, int m,
                                      int clear, int force_cuda, exl2b_stream_t stream) {
    (void)force_cuda;
    QMatrix* q = (QMatrix*)h;
    EXL2B_REQUIRE(q && a && c, "null argument");
    EXL2B_REQUIRE(lda >= q->v.K && ldc >= q->v.N, "leading dimensions too small");
    EXL2B_CUDA(cudaSetDevice(q->device));
    if (m == 1 && q->v.layout == LAYOUT_TC && gemv_i8_enabled()) {        // decode row: the HBM-bound integer GEMV (gemv_i8.cu)
        const I8Out o = {q, (half*)c, clear ? 1 : 0};
        const I8Input in = {(const half*)a, nullptr, nullptr, 0.f, I8_PLAIN};
        return gemv_i8_launch(q->device, (cudaStream_t)stream, &o, 1, in);
    }
    // prefill rows: reconstruct + tensor-core GEMM (q_gemm.cu:233-266).  9..16 rows would be two 8-row passes of the tcgen05 kernel: on
    // matrices up to ~20 M weights the dense path is already faster there (4096 x 4096 at 16 rows: 28 us vs 37 us, reference 31 us)
    const bool small_two_pass = m > GEMV_MTOK && (long long)q->v.K * q->v.N <= 20ll * 1000 * 1000;
    if ((m > GEMM_BIG_MIN_ROWS || small_two_pass || !gemm_tc_supported(q->v)) && gemm_big_available())
        return gemm_big_launch(q, (const half*)a, lda, (half*)c, ldc, m, clear ? 1 : 0, (cudaStream_t)stream);
    GemvMat mt = {};
    mt.w = q->v;
    mt.x = (const half*)a;
    mt.ldx = lda;
    mt.c = (half*)c;
    mt.ldc = ldc;
    mt.clear = clear ? 1 : 0;
    return gemv_launch(q->device, (cudaStream_t)stream, &mt, 1, m, nullptr, 0.f, EPI_STORE);
}

extern "C" int exl2b_gemm_half_q_half_host(exl2b_qmatrix_t h, const uint16_t* a_host, uint16_t* c_host, int m,
                                           exl2b_stream_t stream_) {
    QMatrix* q = (QMatrix*)h;
    EXL2B_REQUIRE(q && a_host && c_host && m > 0, "bad argument");
    cudaStream_t stream = (cudaStream_t)stream_;
    EXL2B_CUDA(cudaSetDevice(q->device));
    const size_t ab = (size_t)m * q->v.K * 2, cb = (size_t)m * q->v.N * 2;
    uint16_t *da = nullptr, *dc = nullptr;
    EXL2B_CUDA(cudaMallocAsync(&da, ab, stream));
    EXL2B_CUDA(cudaMallocAsync(&dc, cb, stream));
    EXL2B_CUDA(cudaMemcpyAsync(da, a_host, ab, cudaMemcpyHostToDevice, stream));
    int rc = exl2b_gemm_half_q_half(h, da, q->v.K, dc, q->v.N, m, 1, 0, stream_);
    if (rc == 0) {
        EXL2B_CUDA(cudaMemcpyAsync(c_host, dc, cb, cudaMemcpyDeviceToHost, stream));
    }
    cudaFreeAsync(da, stream);
    cudaFreeAsync(dc, stream);
    EXL2B_CUDA(cudaStreamSynchronize(stream));
    return rc;
}

// ---- tuning / diagnostics hooks (not part of the reference surface) ---------------------------------------------------
extern "C" int exl2b_debug_set(int ctas_per_sm, unsigned long long* stamps, int cta) {
    if (ctas_per_sm > 0) { exl2b::g_ctas_per_sm = ctas_per_sm; exl2b::g_tc_ctas_per_sm = ctas_per_sm; }
    exl2b::g_dbg = stamps;
    exl2b::g_dbg_cta = cta;
    exl2b::g_dbg_slot = 0;
    return 0;
}
extern "C" int exl2b_debug_set_records(unsigned long long* records) {
    exl2b::g_dbg_rec = records;
    return 0;
}
