// Host emulation of the private layout + fp16-domain unpack + mma fragment mapping (no GPU needed).
// Built and run by tests/test_layout_emu.py with g++.  Exit code 0 = all checks passed.
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>
#include <random>
#include "../../exllamav2_b200/csrc/dequant.cuh"

using namespace exl2b;

static int fails = 0;
#define CHECK(c, ...) do { if (!(c)) { if (fails < 20) { printf("FAIL %s:%d: ", __FILE__, __LINE__); printf(__VA_ARGS__); printf("\n"); } fails++; } } while (0)

// PTX m16n8k16 fragment maps (row-major A 16x16, col B 16x8, C 16x8), lane = g*4 + t
static void a_frag_pos(int lane, int reg, int e, int& row, int& col) {
    int g = lane >> 2, t = lane & 3;
    row = g + ((reg & 1) ? 8 : 0);
    col = t * 2 + e + ((reg & 2) ? 8 : 0);
}
static void b_frag_pos(int lane, int reg, int e, int& k, int& n) {
    int g = lane >> 2, t = lane & 3;
    k = t * 2 + e + (reg ? 8 : 0);
    n = g;
}

template <int BITS> static void test_bits(std::mt19937& rng) {
    constexpr int Pm = plane_main(BITS), Pe = plane_extra(BITS);
    const int zp = 1 << (BITS - 1);
    // constants must be exactly representable
    for (int p = 0; p < 16; ++p) {
        int c = Pe ? PairConst<BITS>::k_double(p, zp) : PairConst<BITS>::c_single(p, zp);
        double back = h_to_d(f16_bits_of_int(c));
        CHECK(back == (double)c, "bits %d pair %d const %d not exact (%f)", BITS, p, c, back);
    }
    for (int trial = 0; trial < 64; ++trial) {
        // a full block: q[n_local][k_local]
        uint32_t q[32][32];
        for (auto& r : q) for (auto& v : r) v = rng() & ((1u << BITS) - 1u);
        if (trial == 0) for (auto& r : q) for (auto& v : r) v = (1u << BITS) - 1u;
        if (trial == 1) for (auto& r : q) for (auto& v : r) v = 0;
        std::vector<uint32_t> block(32 * BITS, 0u);
        for (int lane = 0; lane < 32; ++lane) {
            uint32_t vals[32], mw[8] = {0}, ew[4] = {0};
            for (int i = 0; i < 32; ++i) { ValuePos vp = value_pos(lane, i); vals[i] = q[vp.n_local][vp.k_local];
                CHECK(lane_of(vp.n_local, vp.k_local) == lane && index_of(vp.n_local, vp.k_local) == i, "inverse map"); }
            compose_lane_words(BITS, vals, mw, ew);
            for (int w = 0; w < Pm; ++w) block[main_word_index(BITS, lane, w)] = mw[w];
            for (int w = 0; w < Pe; ++w) block[extra_word_index(BITS, lane, w)] = ew[w];
        }
        // every word index written exactly once is implied by size; now unpack
        float W[32][32];   // dequantised (q - zp) by (n_local, k_local) via the A-fragment route
        bool seen[32][32] = {};
        for (int lane = 0; lane < 32; ++lane) {
            uint32_t mw[8], ew[4] = {0}, A[16];
            for (int w = 0; w < Pm; ++w) mw[w] = block[main_word_index(BITS, lane, w)];
            for (int w = 0; w < Pe; ++w) ew[w] = block[extra_word_index(BITS, lane, w)];
            for (int i = 0; i < 32; ++i) CHECK(extract_value(BITS, mw, ew, i) == q[value_pos(lane, i).n_local][value_pos(lane, i).k_local], "int extract");
            dequant_block_exl2<BITS>(mw, ew, A);
            for (int sub = 0; sub < 2; ++sub) for (int s = 0; s < 2; ++s) for (int reg = 0; reg < 4; ++reg) for (int e = 0; e < 2; ++e) {
                int row, col; a_frag_pos(lane, reg, e, row, col);
                uint32_t bitsv = A[(sub * 2 + s) * 4 + reg];
                double v = h_to_d((uint16_t)(e ? (bitsv >> 16) : bitsv));
                // mma tile (sub, s): row -> n_local = sub*16 + row ; col -> which k?  defined through the B map:
                // B frag of step s: reg0 = act[8t+4s+{0,1}], reg1 = act[8t+4s+{2,3}] -> k index of mma col c:
                // col c = 2t'+e' (+8) with t' = lane&3 of the B lane; A col uses the same t (same lane&3).
                int t = lane & 3;
                int k_local = 8 * t + 4 * s + ((reg & 2) ? 2 : 0) + e;
                (void)col;
                int n_local = sub * 16 + row;
                CHECK(!seen[n_local][k_local], "duplicate (%d,%d)", n_local, k_local);
                seen[n_local][k_local] = true;
                W[n_local][k_local] = (float)v;
                CHECK(v == (double)((int)q[n_local][k_local] - zp), "bits %d lane %d val mismatch n%d k%d got %f want %d", BITS, lane, n_local, k_local, v, (int)q[n_local][k_local] - zp);
            }
        }
        for (int n = 0; n < 32; ++n) for (int k = 0; k < 32; ++k) CHECK(seen[n][k], "unseen (%d,%d)", n, k);
        // consistency of A col <-> B row: mma sums over col c: A[row][c] * B[c][n]; A col c held by lane with t = (c%8)/2,
        // B row c held by lanes with the same t -> both map c -> k_local = 8t + 4s + 2*(c>=8) + (c&1): check bijection
        for (int s = 0; s < 2; ++s) { bool ks[32] = {}; for (int c = 0; c < 16; ++c) { int t = (c & 7) >> 1; int kl = 8 * t + 4 * s + ((c >= 8) ? 2 : 0) + (c & 1); ks[kl] = true; }
            int cnt = 0; for (bool b : ks) cnt += b; CHECK(cnt == 16, "k coverage"); }
    }
}

int main() {
    std::mt19937 rng(1234);
    test_bits<2>(rng); test_bits<3>(rng); test_bits<4>(rng); test_bits<5>(rng); test_bits<6>(rng); test_bits<8>(rng);
    // GPTQ per-row zero
    for (int z1 = 1; z1 <= 16; ++z1) for (int rr = 0; rr < 2; ++rr) {
        uint32_t vals[32], mw[8], ew[4], A[16], zc[4];
        for (int i = 0; i < 32; ++i) vals[i] = rng() & 15;
        compose_lane_words(4, vals, mw, ew);
        for (int j = 0; j < 4; ++j) zc[j] = gptq_zero_const(j & 1, z1);
        dequant_block_gptq(mw, zc, A);
        for (int p = 0; p < 16; ++p) for (int e = 0; e < 2; ++e) {
            double v = h_to_d((uint16_t)(e ? (A[p] >> 16) : A[p]));
            CHECK(v == (double)((int)vals[p * 2 + e] - z1), "gptq z1 %d", z1);
        }
    }
    // 4-bit offset form: A[p] == 64 + q exactly
    for (int trial = 0; trial < 64; ++trial) {
        uint32_t vals[32], mw[8], ew[4], A[16];
        for (int i = 0; i < 32; ++i) vals[i] = rng() & 15;
        compose_lane_words(4, vals, mw, ew);
        dequant_block_4bit_offset(mw, A);
        for (int p = 0; p < 16; ++p) for (int e = 0; e < 2; ++e) {
            double v = h_to_d((uint16_t)(e ? (A[p] >> 16) : A[p]));
            CHECK(v == (double)(OFFSET4 + (int)vals[p * 2 + e]), "offset form p %d e %d got %f", p, e, v);
        }
    }
    for (int trial = 0; trial < 64; ++trial) {
        uint32_t vals[32], mw[8], ew[4], A[16];
        for (int i = 0; i < 32; ++i) vals[i] = rng() & 15;
        compose_lane_words(4, vals, mw, ew);
        dequant_block_4bit_offset2(mw, A);
        for (int p = 0; p < 16; ++p) for (int e = 0; e < 2; ++e) {
            double v = h_to_d((uint16_t)(e ? (A[p] >> 16) : A[p]));
            CHECK(v == (double)(offset2_of_pair(p) + (int)vals[p * 2 + e]), "offset2 form p %d e %d got %f", p, e, v);
        }
    }
    (void)b_frag_pos;
    if (fails) { printf("%d checks failed\n", fails); return 1; }
    printf("layout emulation OK\n");
    return 0;
}
