// Fused attention / MLP blocks behind make_q_attn / q_attn_forward_1 / q_attn_forward_2 (ext_qattn.cpp:24-191,
// cuda/q_attn.cu:153-345) and make_q_mlp / q_mlp_forward_ (ext_qmlp.cpp:22-118, cuda/q_mlp.cu:78-236).
//
// Kernel count per decoder layer (decode):   reference            here
//   attn part 1   norm, Q, K, V, rope        5 launches           2  (norm+QKV in one GEMV launch, rope)
//   attn part 2   O (+residual via atomics)  1                    1  (residual add in the epilogue)
//   mlp           norm, gate, up, act, down  5                    2  (norm+gate|up+silu*mul, down+residual)
// All launches carry the programmatic-dependent-launch attribute, so a whole decode step captured in one CUDA
// graph (exllamav2_b200/model.py) streams weights back to back.
#include "gemv.cuh"
#include "gemv_i8.cuh"

namespace exl2b {

int rope_launch(cudaStream_t stream, half* x, const half* sin, const half* cos, int batch, int rows_per_batch, int head_dim,
                int num_heads, int past_len, const int32_t* past_lens, int neox, int sincos_size);

struct QAttn {
    exl2b_qattn_desc d;
    int device;
    bool i8_qkv;          // q/k/v share K and the row permutation: one gemv_i8 launch for a single row
};
struct QMlp {
    exl2b_qmlp_desc d;
    int device;
    bool i8_gu;           // same for gate/up
    half* up_scratch;     // single-row up projection when the caller passes no temp_b (reference: temp_b of make_q_mlp)
};

static GemvMat make_mat(const QMatrix* q, const half* x, int ldx, half* c, int ldc, int clear) {
    GemvMat m = {};
    m.w = q->v;
    m.x = x;
    m.ldx = ldx;
    m.c = c;
    m.ldc = ldc;
    m.clear = clear;
    return m;
}

// epilogue of a producer launch -> the consumers' activation buffers (+ sums of squares when they apply an RMSNorm)
static int chain_out(GemvExtras& ex, const exl2b_chain_t* next) {
    if (!next || next->num_consumers <= 0) return 0;
    EXL2B_REQUIRE(next->num_consumers <= GEMV_MAX_MATS, "at most %d chained consumers", GEMV_MAX_MATS);
    for (int i = 0; i < next->num_consumers; ++i) {
        QMatrix* c = (QMatrix*)next->consumers[i];
        EXL2B_REQUIRE(c && c->v.layout == LAYOUT_TC, "chained consumer must be a tcgen05-layout matrix");
        int rc = qmatrix_chain_buffers(c);
        if (rc) return rc;
        ex.scat[i] = ScatterTarget{c->xp_buf, c->invperm, (const half*)next->norm_weight};
    }
    ex.num_scat = next->num_consumers;
    if (next->norm_weight) ex.sumsq_out = ((QMatrix*)next->consumers[0])->sumsq_buf;
    return 0;
}
// consumer side: the matrices' inputs were written by a chained producer
static int chain_in(GemvExtras& ex, GemvMat* mats, const QMatrix* const* qs, int nm, bool has_norm) {
    ex.prepared = 1;
    for (int i = 0; i < nm; ++i) {
        EXL2B_REQUIRE(qs[i]->xp_buf, "input_prepared set, but no chained producer has written this matrix's input");
        mats[i].xp = qs[i]->xp_buf;
    }
    if (has_norm) {
        ex.sumsq_in = qs[0]->sumsq_buf;
        ex.sumsq_in_strips = (qs[0]->v.K + 127) / 128;
    }
    return 0;
}

// i8 (single-row) form of a chain: the producer's finalisation scatters a second fp16 copy of its output into the first
// consumer's row buffer, in that matrix's stored-row order (all consumers of one chain share their permutation)
static int chain_out_i8(I8Out& o, const exl2b_chain_t* next, int slot = 0) {
    if (!next || next->num_consumers <= 0) return 0;
    QMatrix* c = (QMatrix*)next->consumers[0];
    EXL2B_REQUIRE(c && c->v.layout == LAYOUT_TC, "chained consumer must be a default-layout matrix");
    int rc = qmatrix_chain_buffers(c);
    if (rc) return rc;
    o.c_perm = c->xp_buf + (size_t)slot * c->v.K;
    o.out_invperm = c->invperm;
    return 0;
}

}  // namespace exl2b

using namespace exl2b;

extern "C" int exl2b_qattn_create(const exl2b_qattn_desc* d, exl2b_qattn_t* out) {
    EXL2B_REQUIRE(d && out, "null argument");
    // o_proj may be absent: a tensor-parallel rank runs part 1 on its heads and applies its column shard of o_proj itself
    // (exllamav2_b200/tensor_p.py), part 2 then is not available on this handle
    EXL2B_REQUIRE(d->q_proj && d->k_proj && d->v_proj, "q/k/v handles are required");
    const QMatrix *q = (const QMatrix*)d->q_proj, *k = (const QMatrix*)d->k_proj, *v = (const QMatrix*)d->v_proj,
                  *o = d->o_proj ? (const QMatrix*)d->o_proj : q;
    EXL2B_REQUIRE(q->v.K == d->hidden_size && k->v.K == d->hidden_size && v->v.K == d->hidden_size, "q/k/v_proj is wrong shape");
    EXL2B_REQUIRE(!d->o_proj || o->v.N == d->hidden_size, "o_proj is wrong shape");          // ext_qattn.cpp:67
    EXL2B_REQUIRE(q->v.N == d->num_heads * d->head_dim && k->v.N == d->num_kv_heads * d->head_dim && v->v.N == k->v.N,
                  "projection widths do not match the head layout");
    EXL2B_REQUIRE(q->device == k->device && q->device == v->device && q->device == o->device, "handles on different devices");
    const QMatrix* qkv[3] = {q, k, v};
    QAttn* a = new QAttn{*d, q->device, gemv_i8_fusable(qkv, 3)};
    *out = (exl2b_qattn_t)a;
    return 0;
}

extern "C" int exl2b_qattn_destroy(exl2b_qattn_t h) {
    delete (QAttn*)h;
    return 0;
}

extern "C" int exl2b_qattn_forward_1_ex(exl2b_qattn_t h, const uint16_t* x, int batch, int q_len, int past_len,
                                        const int32_t* past_lens, uint16_t* q, uint16_t* k, uint16_t* v, const uint16_t* sin,
                                        const uint16_t* cos, int input_prepared, exl2b_stream_t stream_) {
    QAttn* a = (QAttn*)h;
    EXL2B_REQUIRE(a && q && k && v && (x || input_prepared), "null argument");
    cudaStream_t stream = (cudaStream_t)stream_;
    EXL2B_CUDA(cudaSetDevice(a->device));
    const exl2b_qattn_desc& d = a->d;
    const int rows = batch * q_len;
    const QMatrix *mq = (const QMatrix*)d.q_proj, *mk = (const QMatrix*)d.k_proj, *mv = (const QMatrix*)d.v_proj;
    GemvMat mats[3] = {
        make_mat(mq, (const half*)x, d.hidden_size, (half*)q, mq->v.N, 1),
        make_mat(mk, (const half*)x, d.hidden_size, (half*)k, mk->v.N, 1),
        make_mat(mv, (const half*)x, d.hidden_size, (half*)v, mv->v.N, 1),
    };
    const bool rope = d.rope_style != 0;
    if (rows == 1 && a->i8_qkv && gemv_i8_enabled()) {
        // decode row: RMSNorm is the GEMV's prologue, Q|K|V are one launch (gemv_i8.cu).  input_prepared: the row was left in
        // q_proj's stored-row order by the producer launch (I8Out::c_perm -> q_proj's row buffer).  RoPE as in the reference
        // (q_attn.cu:271-300) unless the caller passes no tables: exl2b_paged_attn_decode_q4_ex rotates q / k as it reads them.
        const I8Out o[3] = {{mq, (half*)q, 1}, {mk, (half*)k, 1}, {mv, (half*)v, 1}};
        I8Input in = {(const half*)x, nullptr, (const half*)d.layernorm, d.norm_epsilon, d.layernorm ? I8_RMSNORM : I8_PLAIN, 0};
        if (input_prepared) {
            EXL2B_REQUIRE(mq->xp_buf, "input_prepared set, but no chained producer has written q_proj's input row");
            in.x = mq->xp_buf;
            in.x_permuted = 1;
        }
        int rc = gemv_i8_launch(a->device, stream, o, 3, in);
        if (rc || !rope || !sin) return rc;
        const int neox = d.rope_style == 2;
        rc = rope_launch(stream, (half*)q, (const half*)sin, (const half*)cos, batch, q_len * d.num_heads, d.head_dim, d.num_heads,
                         past_len, past_lens, neox, d.sincos_size);
        if (rc) return rc;
        return rope_launch(stream, (half*)k, (const half*)sin, (const half*)cos, batch, q_len * d.num_kv_heads, d.head_dim,
                           d.num_kv_heads, past_len, past_lens, neox, d.sincos_size);
    }
    if (rope) EXL2B_REQUIRE(sin && cos, "rope needs sin/cos tables");
    if (rows > GEMM_BIG_MIN_ROWS && !input_prepared && gemm_big_available()) {
        // prefill rows: the reference's own sequence (q_attn.cu:153-300) -- rms_norm, three GEMMs, rope -- with the GEMMs on the
        // tensor cores (gemm_big.cu)
        const half* xin = (const half*)x;
        half* xn = nullptr;
        if (d.layernorm) {
            EXL2B_CUDA(cudaMallocAsync(&xn, (size_t)rows * d.hidden_size * sizeof(half), stream));
            int rc = exl2b_rms_norm(x, d.layernorm, (uint16_t*)xn, d.norm_epsilon, rows, d.hidden_size, stream_);
            if (rc) return rc;
            xin = xn;
        }
        int rc = gemm_big_launch(mq, xin, d.hidden_size, (half*)q, mq->v.N, rows, 1, stream);
        if (!rc) rc = gemm_big_launch(mk, xin, d.hidden_size, (half*)k, mk->v.N, rows, 1, stream);
        if (!rc) rc = gemm_big_launch(mv, xin, d.hidden_size, (half*)v, mv->v.N, rows, 1, stream);
        if (xn) cudaFreeAsync(xn, stream);
        if (rc || !rope) return rc;
        const int neox = d.rope_style == 2;
        rc = rope_launch(stream, (half*)q, (const half*)sin, (const half*)cos, batch, q_len * d.num_heads, d.head_dim, d.num_heads,
                         past_len, past_lens, neox, d.sincos_size);
        if (rc) return rc;
        return rope_launch(stream, (half*)k, (const half*)sin, (const half*)cos, batch, q_len * d.num_kv_heads, d.head_dim,
                           d.num_kv_heads, past_len, past_lens, neox, d.sincos_size);
    }
    const bool fuse = gemv_supports_extras(mats, 3, rows) && (!rope || (d.head_dim <= 128 && 128 % d.head_dim == 0 && d.sincos_size <= d.head_dim));
    EXL2B_REQUIRE(!input_prepared || fuse, "input_prepared needs the tcgen05 layout and at most %d rows", GEMV_MTOK);
    if (fuse) {
        GemvExtras ex = {};
        if (rope) ex.rope = RopeFuse{(const half*)sin, (const half*)cos, past_lens, past_len, q_len, d.head_dim, d.sincos_size, d.rope_style == 2, 3u};
        if (input_prepared) {
            const QMatrix* qs[3] = {mq, mk, mv};
            int rc = chain_in(ex, mats, qs, 3, d.layernorm != nullptr);
            if (rc) return rc;
        }
        return gemv_launch(a->device, stream, mats, 3, rows, (const half*)d.layernorm, d.norm_epsilon, EPI_STORE, &ex);
    }
    int rc = gemv_launch(a->device, stream, mats, 3, rows, (const half*)d.layernorm, d.norm_epsilon, EPI_STORE);
    if (rc) return rc;
    if (rope) {
        const int neox = d.rope_style == 2;
        rc = rope_launch(stream, (half*)q, (const half*)sin, (const half*)cos, batch, q_len * d.num_heads, d.head_dim, d.num_heads,
                         past_len, past_lens, neox, d.sincos_size);
        if (rc) return rc;
        rc = rope_launch(stream, (half*)k, (const half*)sin, (const half*)cos, batch, q_len * d.num_kv_heads, d.head_dim,
                         d.num_kv_heads, past_len, past_lens, neox, d.sincos_size);
    }
    return rc;
}

extern "C" int exl2b_qattn_forward_1(exl2b_qattn_t h, const uint16_t* x, int batch, int q_len, int past_len,
                                     const int32_t* past_lens, uint16_t* q, uint16_t* k, uint16_t* v, const uint16_t* sin,
                                     const uint16_t* cos, exl2b_stream_t stream) {
    EXL2B_REQUIRE(x, "null argument");
    return exl2b_qattn_forward_1_ex(h, x, batch, q_len, past_len, past_lens, q, k, v, sin, cos, 0, stream);
}

extern "C" int exl2b_qattn_forward_2_ex(exl2b_qattn_t h, uint16_t* x, const uint16_t* attn_out, int batch, int q_len,
                                        int input_prepared, const exl2b_chain_t* next, exl2b_stream_t stream) {
    QAttn* a = (QAttn*)h;
    EXL2B_REQUIRE(a && x && (attn_out || input_prepared), "null argument");
    EXL2B_REQUIRE(a->d.o_proj, "this attention handle was created without o_proj");
    EXL2B_CUDA(cudaSetDevice(a->device));
    const QMatrix* mo = (const QMatrix*)a->d.o_proj;
    GemvMat m = make_mat(mo, (const half*)attn_out, mo->v.K, (half*)x, mo->v.N, a->d.has_residual ? 0 : 1);
    const bool want = input_prepared || (next && next->num_consumers > 0);
    if (batch * q_len == 1 && mo->v.layout == LAYOUT_TC && gemv_i8_enabled()) {
        I8Out o = {mo, (half*)x, a->d.has_residual ? 0 : 1};
        I8Input in = {(const half*)attn_out, nullptr, nullptr, 0.f, I8_PLAIN, 0};
        if (input_prepared) {
            EXL2B_REQUIRE(mo->xp_buf, "input_prepared set, but no chained producer has written o_proj's input row");
            in.x = mo->xp_buf;
            in.x_permuted = 1;
        }
        int rc = chain_out_i8(o, next);
        if (rc) return rc;
        return gemv_i8_launch(a->device, (cudaStream_t)stream, &o, 1, in);
    }
    if (!want && batch * q_len > GEMM_BIG_MIN_ROWS && gemm_big_available())
        return gemm_big_launch(mo, (const half*)attn_out, mo->v.K, (half*)x, mo->v.N, batch * q_len, a->d.has_residual ? 0 : 1, (cudaStream_t)stream);
    if (!want) return gemv_launch(a->device, (cudaStream_t)stream, &m, 1, batch * q_len, nullptr, 0.f, EPI_STORE);
    EXL2B_REQUIRE(gemv_supports_extras(&m, 1, batch * q_len), "chained launches need the tcgen05 layout and at most %d rows", GEMV_MTOK);
    GemvExtras ex = {};
    int rc = chain_out(ex, next);
    if (rc) return rc;
    if (input_prepared) {
        rc = chain_in(ex, &m, &mo, 1, false);
        if (rc) return rc;
    }
    return gemv_launch(a->device, (cudaStream_t)stream, &m, 1, batch * q_len, nullptr, 0.f, EPI_STORE, &ex);
}

extern "C" int exl2b_qattn_forward_2(exl2b_qattn_t h, uint16_t* x, const uint16_t* attn_out, int batch, int q_len,
                                     exl2b_stream_t stream) {
    EXL2B_REQUIRE(attn_out, "null argument");
    return exl2b_qattn_forward_2_ex(h, x, attn_out, batch, q_len, 0, nullptr, stream);
}

extern "C" int exl2b_qmlp_create(const exl2b_qmlp_desc* d, exl2b_qmlp_t* out) {
    EXL2B_REQUIRE(d && out, "null argument");
    // down may be absent (tensor-parallel rank: gate|up on its intermediate slice, exl2b_qmlp_forward_gateup)
    EXL2B_REQUIRE(d->gate && d->up, "gate/up handles are required");
    const QMatrix *g = (const QMatrix*)d->gate, *u = (const QMatrix*)d->up, *dn = d->down ? (const QMatrix*)d->down : nullptr;
    EXL2B_REQUIRE(g->v.K == d->hidden_size && u->v.K == d->hidden_size && (!dn || dn->v.N == d->hidden_size), "mlp matrices have wrong shape");
    EXL2B_REQUIRE(g->v.N == d->intermediate_size && u->v.N == d->intermediate_size && (!dn || dn->v.K == d->intermediate_size),
                  "mlp intermediate size mismatch");
    EXL2B_REQUIRE(g->device == u->device && (!dn || g->device == dn->device), "handles on different devices");
    const QMatrix* gu[2] = {g, u};
    QMlp* m = new QMlp{*d, g->device, gemv_i8_fusable(gu, 2), nullptr};
    *out = (exl2b_qmlp_t)m;
    return 0;
}

extern "C" int exl2b_qmlp_destroy(exl2b_qmlp_t h) {
    QMlp* m = (QMlp*)h;
    if (m && m->up_scratch) {
        cudaSetDevice(m->device);
        cudaFree(m->up_scratch);
    }
    delete m;
    return 0;
}

extern "C" int exl2b_qmlp_forward_ex(exl2b_qmlp_t h, uint16_t* x, int rows, uint16_t* temp_a, uint16_t* temp_b,
                                     int input_prepared, const exl2b_chain_t* next, exl2b_stream_t stream_) {
    (void)temp_b;    // the up projection never materialises: silu(gate)*up is formed in the GEMV epilogue
    QMlp* m = (QMlp*)h;
    EXL2B_REQUIRE(m && x && temp_a, "null argument");
    EXL2B_REQUIRE(m->d.down, "this MLP handle was created without down_proj");
    cudaStream_t stream = (cudaStream_t)stream_;
    EXL2B_CUDA(cudaSetDevice(m->device));
    const exl2b_qmlp_desc& d = m->d;
    const QMatrix *g = (const QMatrix*)d.gate, *u = (const QMatrix*)d.up, *dn = (const QMatrix*)d.down;
    GemvMat gu[2] = {
        make_mat(g, (const half*)x, d.hidden_size, (half*)temp_a, d.intermediate_size, 1),
        make_mat(u, (const half*)x, d.hidden_size, (half*)temp_a, d.intermediate_size, 1),
    };
    GemvMat down = make_mat(dn, (const half*)temp_a, d.intermediate_size, (half*)x, d.hidden_size, d.has_residual ? 0 : 1);
    if (rows == 1 && m->i8_gu && dn->v.layout == LAYOUT_TC && gemv_i8_enabled()) {
        // decode row: gate|up in one launch with RMSNorm as its prologue; act(gate) * up is the PROLOGUE of the down launch,
        // which reads both rows in its own stored-row order (scattered there by the gate|up launch's finalisation)
        half* tb = (half*)temp_b;
        if (!tb) {
            if (!m->up_scratch) EXL2B_CUDA(cudaMalloc(&m->up_scratch, (size_t)d.intermediate_size * sizeof(half)));
            tb = m->up_scratch;
        }
        int rc = qmatrix_chain_buffers(const_cast<QMatrix*>(dn));
        if (rc) return rc;
        I8Out o[2] = {{g, (half*)temp_a, 1}, {u, tb, 1}};
        o[0].c_perm = dn->xp_buf;
        o[1].c_perm = dn->xp_buf + dn->v.K;
        o[0].out_invperm = o[1].out_invperm = dn->invperm;
        I8Input in1 = {(const half*)x, nullptr, (const half*)d.layernorm, d.norm_epsilon, d.layernorm ? I8_RMSNORM : I8_PLAIN, 0};
        if (input_prepared) {
            EXL2B_REQUIRE(g->xp_buf, "input_prepared set, but no chained producer has written gate_proj's input row");
            in1.x = g->xp_buf;
            in1.x_permuted = 1;
        }
        rc = gemv_i8_launch(m->device, stream, o, 2, in1);
        if (rc) return rc;
        I8Out od = {dn, (half*)x, d.has_residual ? 0 : 1};
        rc = chain_out_i8(od, next);
        if (rc) return rc;
        const I8Input in2 = {dn->xp_buf, dn->xp_buf + dn->v.K, nullptr, 0.f, d.act_gelu ? I8_GELU_MUL : I8_SILU_MUL, 1};
        return gemv_i8_launch(m->device, stream, &od, 1, in2);
    }
    if (rows > GEMM_BIG_MIN_ROWS && !input_prepared && !(next && next->num_consumers > 0) && gemm_big_available()) {
        // prefill rows: rms_norm, gate and up GEMMs, act*mul, down GEMM (+residual) as in q_mlp.cu:78-236, GEMMs on the tensor cores
        const half* xin = (const half*)x;
        half *xn = nullptr, *tb = (half*)temp_b;
        bool own_tb = false;
        if (d.layernorm) {
            EXL2B_CUDA(cudaMallocAsync(&xn, (size_t)rows * d.hidden_size * sizeof(half), stream));
            int rc = exl2b_rms_norm(x, d.layernorm, (uint16_t*)xn, d.norm_epsilon, rows, d.hidden_size, stream_);
            if (rc) return rc;
            xin = xn;
        }
        if (!tb) {
            EXL2B_CUDA(cudaMallocAsync(&tb, (size_t)rows * d.intermediate_size * sizeof(half), stream));
            own_tb = true;
        }
        int rc = gemm_big_launch(g, xin, d.hidden_size, (half*)temp_a, d.intermediate_size, rows, 1, stream);
        if (!rc) rc = gemm_big_launch(u, xin, d.hidden_size, tb, d.intermediate_size, rows, 1, stream);
        if (!rc) rc = exl2b_act_mul(temp_a, (const uint16_t*)tb, rows, d.intermediate_size, d.act_gelu, stream_);
        if (!rc) rc = gemm_big_launch(dn, (const half*)temp_a, d.intermediate_size, (half*)x, d.hidden_size, rows, d.has_residual ? 0 : 1, stream);
        if (xn) cudaFreeAsync(xn, stream);
        if (own_tb) cudaFreeAsync(tb, stream);
        return rc;
    }
    const int epi = d.act_gelu ? EPI_GELU_MUL : EPI_SILU_MUL;
    const bool fuse = gemv_supports_extras(gu, 2, rows) && gemv_supports_extras(&down, 1, rows);
    EXL2B_REQUIRE(fuse || (!input_prepared && !(next && next->num_consumers > 0)),
                  "chained launches need the tcgen05 layout and at most %d rows", GEMV_MTOK);
    if (!fuse) {
        int rc = gemv_launch(m->device, stream, gu, 2, rows, (const half*)d.layernorm, d.norm_epsilon, epi);
        if (rc) return rc;
        return gemv_launch(m->device, stream, &down, 1, rows, nullptr, 0.f, EPI_STORE);
    }
    // gate|up writes silu(gate)*up straight into down's activation buffer (permuted, UMMA layout): no prep launch between
    GemvExtras e1 = {};
    exl2b_chain_t to_down = {};
    to_down.consumers[0] = (exl2b_qmatrix_t)dn;
    to_down.num_consumers = 1;
    int rc = chain_out(e1, &to_down);
    if (rc) return rc;
    if (input_prepared) {
        const QMatrix* qs[2] = {g, u};
        rc = chain_in(e1, gu, qs, 2, d.layernorm != nullptr);
        if (rc) return rc;
    }
    rc = gemv_launch(m->device, stream, gu, 2, rows, (const half*)d.layernorm, d.norm_epsilon, epi, &e1);
    if (rc) return rc;
    GemvExtras e2 = {};
    rc = chain_out(e2, next);
    if (rc) return rc;
    rc = chain_in(e2, &down, &dn, 1, false);
    if (rc) return rc;
    return gemv_launch(m->device, stream, &down, 1, rows, nullptr, 0.f, EPI_STORE, &e2);
}

// first half of the MLP only: temp_a[rows, intermediate] = act(norm(x) @ gate) * (norm(x) @ up)   (x is not modified)
extern "C" int exl2b_qmlp_forward_gateup(exl2b_qmlp_t h, const uint16_t* x, int rows, uint16_t* temp_a, exl2b_stream_t stream_) {
    QMlp* m = (QMlp*)h;
    EXL2B_REQUIRE(m && x && temp_a, "null argument");
    EXL2B_CUDA(cudaSetDevice(m->device));
    const exl2b_qmlp_desc& d = m->d;
    const QMatrix *g = (const QMatrix*)d.gate, *u = (const QMatrix*)d.up;
    if (rows == 1 && m->i8_gu && gemv_i8_enabled()) {
        // decode row: gate|up as one integer-GEMV launch (RMSNorm in its prologue), then act(gate) * up on the rank's slice
        if (!m->up_scratch) EXL2B_CUDA(cudaMalloc(&m->up_scratch, (size_t)d.intermediate_size * sizeof(half)));
        const I8Out o[2] = {{g, (half*)temp_a, 1}, {u, m->up_scratch, 1}};
        const I8Input in = {(const half*)x, nullptr, (const half*)d.layernorm, d.norm_epsilon, d.layernorm ? I8_RMSNORM : I8_PLAIN, 0};
        int rc = gemv_i8_launch(m->device, (cudaStream_t)stream_, o, 2, in);
        if (rc) return rc;
        return exl2b_act_mul(temp_a, (const uint16_t*)m->up_scratch, 1, d.intermediate_size, d.act_gelu, stream_);
    }
    GemvMat gu[2] = {
        make_mat(g, (const half*)x, d.hidden_size, (half*)temp_a, d.intermediate_size, 1),
        make_mat(u, (const half*)x, d.hidden_size, (half*)temp_a, d.intermediate_size, 1),
    };
    return gemv_launch(m->device, (cudaStream_t)stream_, gu, 2, rows, (const half*)d.layernorm, d.norm_epsilon,
                       d.act_gelu ? EPI_GELU_MUL : EPI_SILU_MUL);
}

extern "C" int exl2b_qmlp_forward(exl2b_qmlp_t h, uint16_t* x, int rows, uint16_t* temp_a, uint16_t* temp_b,
                                  exl2b_stream_t stream) {
    return exl2b_qmlp_forward_ex(h, x, rows, temp_a, temp_b, 0, nullptr, stream);
}

// gemm_half_q_half whose input was prepared by a chained producer (lm_head after the last MLP: the final RMSNorm is the
// producer's scatter scale + this launch's deferred 1/rms)
extern "C" int exl2b_gemm_half_q_half_prepared(exl2b_qmatrix_t h, uint16_t* c, int ldc, int m, int clear, int has_norm,
                                               float norm_eps, exl2b_stream_t stream) {
    QMatrix* q = (QMatrix*)h;
    EXL2B_REQUIRE(q && c, "null argument");
    EXL2B_REQUIRE(ldc >= q->v.N, "leading dimension too small");
    EXL2B_CUDA(cudaSetDevice(q->device));
    GemvMat mt = make_mat(q, nullptr, q->v.K, (half*)c, ldc, clear ? 1 : 0);
    EXL2B_REQUIRE(gemv_supports_extras(&mt, 1, m), "chained launches need the tcgen05 layout and at most %d rows", GEMV_MTOK);
    GemvExtras ex = {};
    const QMatrix* qc = q;
    int rc = chain_in(ex, &mt, &qc, 1, has_norm != 0);
    if (rc) return rc;
    return gemv_launch(q->device, (cudaStream_t)stream, &mt, 1, m, nullptr, norm_eps, EPI_STORE, &ex);
}

// the activation buffer a chained producer outside this file (the attention kernel) writes for matrix h
extern "C" int exl2b_qmatrix_chain_target(exl2b_qmatrix_t h, uint16_t** xp, const uint16_t** invperm) {
    QMatrix* q = (QMatrix*)h;
    EXL2B_REQUIRE(q && xp && invperm, "null argument");
    EXL2B_REQUIRE(q->v.layout == LAYOUT_TC, "chained consumer must be a tcgen05-layout matrix");
    int rc = qmatrix_chain_buffers(q);
    if (rc) return rc;
    *xp = (uint16_t*)q->xp_buf;
    *invperm = q->invperm;
    return 0;
}

// rms_norm + gemm_half_q_half on ONE row as a single launch (final norm + lm_head of a decode step; the reference runs
// rms_norm_cuda then gemm_half_q_half_cuda, exllamav2/model.py:1036-1044 -> rmsnorm.py:141, linear.py:366).
// x == NULL: the row was left in this matrix's stored-row order by a chained producer launch.
extern "C" int exl2b_gemm_half_q_half_norm(exl2b_qmatrix_t h, const uint16_t* x, const uint16_t* norm_w, float norm_eps,
                                           uint16_t* c, int clear, exl2b_stream_t stream) {
    QMatrix* q = (QMatrix*)h;
    EXL2B_REQUIRE(q && norm_w && c, "null argument");
    EXL2B_REQUIRE(q->v.layout == LAYOUT_TC, "matrix is not in the default layout");
    EXL2B_CUDA(cudaSetDevice(q->device));
    const I8Out o = {q, (half*)c, clear ? 1 : 0};
    I8Input in = {(const half*)x, nullptr, (const half*)norm_w, norm_eps, I8_RMSNORM, 0};
    if (!x) {
        EXL2B_REQUIRE(q->xp_buf, "no input row given and no chained producer has written this matrix's input row");
        in.x = q->xp_buf;
        in.x_permuted = 1;
    }
    return gemv_i8_launch(q->device, (cudaStream_t)stream, &o, 1, in);
}
