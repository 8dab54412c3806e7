"""CPU emulation of the index / byte-order logic of csrc/gemv_i8.cu (no GPU): the parts of the batch-1 integer GEMV
that are pure bookkeeping and that a wrong constant would break silently.

1. consume_slab: for every bit width, compose a 32 k x 32 column block in the tcgen05 layout (layout.h
   compose_lane_words), stage a row of 16-bit integers in the order stage_round writes it (high / low byte planes,
   bytes of octet j ordered k = 8j + {0,4,1,5} | {2,6,3,7}) and run the kernel's mask / shift / PRMT selectors; the
   integer sums must equal sum_k a_k q_k exactly.
2. work split: locate() / CTA ranges / first and last owner of a 64-column pair (split-K workspace slots) for mixed
   bit widths, several matrices per launch and tiny matrices.
"""
import numpy as np
import pytest

# ---- layout.h ---------------------------------------------------------------------------------------------------------


def plane_main(b):
    return {2: 2, 3: 2, 8: 8}.get(b, 4)


def plane_extra(b):
    return {3: 1, 5: 1, 6: 2}.get(b, 0)


def compose_lane_words(bits, q):
    """q: 32 values of one column (k = 0..31) -> (main words, extra words) exactly as layout.h compose_lane_words."""
    Pm, Pe = plane_main(bits), plane_extra(bits)
    mw = [0] * Pm
    ew = [0] * Pe
    for i in range(32):
        p, e = i >> 1, i & 1
        fm = int(q[i]) & ((1 << Pm) - 1)
        ppw = 16 // Pm
        mw[p // ppw] |= fm << (16 * e + Pm * (p % ppw))
        if Pe:
            fe = (int(q[i]) >> Pm) & ((1 << Pe) - 1)
            ppw_e = 16 // Pe
            ew[p // ppw_e] |= fe << (16 * e + Pe * (p % ppw_e))
    return [w & 0xFFFFFFFF for w in mw], [w & 0xFFFFFFFF for w in ew]


# ---- gemv_i8.cu --------------------------------------------------------------------------------------------------------


def byte_perm(a, b, sel):
    src = [(a >> (8 * i)) & 0xFF for i in range(4)] + [(b >> (8 * i)) & 0xFF for i in range(4)]
    out = 0
    for n in range(4):
        out |= src[(sel >> (4 * n)) & 0x7] << (8 * n)
    return out


def dp4a(w, x, signed_x):
    s = 0
    for i in range(4):
        wb = (w >> (8 * i)) & 0xFF
        xb = (x >> (8 * i)) & 0xFF
        if signed_x and xb >= 128:
            xb -= 256
        s += wb * xb
    return s


def stage_row(a):
    """a: 32 int16 values (one slab) -> XH[8], XL[8] as stage_round packs them."""
    XH, XL = [0] * 8, [0] * 8
    for j in range(4):
        q = [int(v) for v in a[8 * j:8 * j + 8]]

        def pack(idx, sh):
            w = 0
            for n, i in enumerate(idx):
                w |= ((q[i] >> sh) & 0xFF) << (8 * n)
            return w
        XH[2 * j], XH[2 * j + 1] = pack((0, 4, 1, 5), 8), pack((2, 6, 3, 7), 8)
        XL[2 * j], XL[2 * j + 1] = pack((0, 4, 1, 5), 0), pack((2, 6, 3, 7), 0)
    return XH, XL


def two_field_operands(XH, XL):
    YH, YL = [0] * 8, [0] * 8
    for w in range(2):
        for i in range(4):
            a, b, hi = 2 * w, 2 * w + 1, i & 1
            sel = 0x7351 if (i & 2) else 0x6240
            YH[w * 4 + i] = byte_perm(XH[2 * a + hi], XH[2 * b + hi], sel)
            YL[w * 4 + i] = byte_perm(XL[2 * a + hi], XL[2 * b + hi], sel)
    return YH, YL


def consume_slab(bits, mw, ew, XH, XL):
    """One column of consume_slab<BITS>; returns the value the flush computes before the zero-point term."""
    Pm, Pe = plane_main(bits), plane_extra(bits)
    am = [0, 0, 0, 0]
    ae = [0, 0]
    if Pm == 4:
        for j in range(4):
            lo, hi = mw[j] & 0x0F0F0F0F, mw[j] & 0xF0F0F0F0
            am[0] += dp4a(lo, XH[2 * j], True)
            am[1] += dp4a(lo, XL[2 * j], False)
            am[2] += dp4a(hi, XH[2 * j + 1], True)
            am[3] += dp4a(hi, XL[2 * j + 1], False)
    elif Pm == 8:
        for w in range(8):
            j = w >> 1
            sel = 0x7351 if (w & 1) else 0x6240
            am[0] += dp4a(mw[w], byte_perm(XH[2 * j], XH[2 * j + 1], sel), True)
            am[1] += dp4a(mw[w], byte_perm(XL[2 * j], XL[2 * j + 1], sel), False)
    else:
        YH, YL = two_field_operands(XH, XL)
        for w in range(2):
            for i in range(4):
                t = (mw[w] >> (2 * i)) & 0x03030303
                am[0] += dp4a(t, YH[w * 4 + i], True)
                am[1] += dp4a(t, YL[w * 4 + i], False)
    if Pe == 1:
        for i in range(8):
            oa, ob, t = i >> 2, 2 + (i >> 2), i & 3
            hi = t & 1
            sel = 0x7351 if (t & 2) else 0x6240
            f = (ew[0] >> i) & 0x01010101
            ae[0] += dp4a(f, byte_perm(XH[2 * oa + hi], XH[2 * ob + hi], sel), True)
            ae[1] += dp4a(f, byte_perm(XL[2 * oa + hi], XL[2 * ob + hi], sel), False)
    elif Pe == 2:
        YH, YL = two_field_operands(XH, XL)
        for w in range(2):
            for i in range(4):
                t = (ew[w] >> (2 * i)) & 0x03030303
                ae[0] += dp4a(t, YH[w * 4 + i], True)
                ae[1] += dp4a(t, YL[w * 4 + i], False)
    hi16 = (am[2] << 8) + am[3]
    assert hi16 % 16 == 0
    return ((am[0] << 8) + am[1]) + (hi16 >> 4) + (((ae[0] << 8) + ae[1]) << Pm)


@pytest.mark.parametrize("bits", [2, 3, 4, 5, 6, 8])
def test_consume_slab_selectors(bits):
    rng = np.random.default_rng(bits)
    for trial in range(20):
        q = rng.integers(0, 1 << bits, size=32)
        a = rng.integers(-32767, 32768, size=32)
        if trial == 0:
            a[:] = 0
            a[trial % 32] = 1
        mw, ew = compose_lane_words(bits, q)
        XH, XL = stage_row(a)
        got = consume_slab(bits, mw, ew, XH, XL)
        assert got == int((a.astype(np.int64) * q.astype(np.int64)).sum())


def test_consume_slab_one_hot_every_k():
    """every k position reaches its own weight (one-hot rows), all bit widths"""
    rng = np.random.default_rng(0)
    for bits in (2, 3, 4, 5, 6, 8):
        q = rng.integers(0, 1 << bits, size=32)
        mw, ew = compose_lane_words(bits, q)
        for k in range(32):
            a = np.zeros(32, dtype=np.int64)
            a[k] = -12345
            XH, XL = stage_row(a)
            assert consume_slab(bits, mw, ew, XH, XL) == -12345 * int(q[k])


# ---- work split ------------------------------------------------------------------------------------------------------------


class Mat:
    def __init__(self, strips, regions, KS):
        """regions: list of (ks_begin, bits); derives off_base like qmatrix.cu build_regions (TC layout: 128*bits per slab)"""
        self.strips, self.KS = strips, KS
        self.reg = []
        off = 0
        for i, (ks0, bits) in enumerate(regions):
            self.reg.append((ks0, bits, off))
            ks1 = regions[i + 1][0] if i + 1 < len(regions) else KS
            off += (ks1 - ks0) * 128 * bits
        self.blk_stream_bytes = off
        self.pairs = strips * 2


def locate(mats, KS, pos):
    mi = 0
    for i, m in enumerate(mats):
        if i and pos >= m.byte_base:
            mi = i
    m = mats[mi]
    rel = pos - m.byte_base
    pair_bytes = 2 * m.blk_stream_bytes
    p = rel // pair_bytes
    off = (rel - p * pair_bytes) >> 1
    r = 0
    for i, (ks0, bits, ob) in enumerate(m.reg):
        if i and off >= ob:
            r = i
    ks0, bits, ob = m.reg[r]
    return (m.gp_base + p) * KS + ks0 + (off - ob) // (128 * bits)


def split(mats, KS, sms=148):
    bytes_, gp, max_bits = 0, 0, 2
    for m in mats:
        m.byte_base, m.gp_base = bytes_, gp
        gp += m.pairs
        bytes_ += m.strips * 4 * m.blk_stream_bytes
        max_bits = max(max_bits, max(b for _, b, _ in m.reg))
    C = max(1, min(sms, bytes_ // (256 * max_bits)))
    starts = [locate(mats, KS, bytes_ * c // C) for c in range(C)] + [gp * KS]
    return bytes_, gp, C, starts


def owners(mats, KS, B, C, gp):
    mi = 0
    for i, m in enumerate(mats):
        if i and gp >= m.gp_base:
            mi = i
    m = mats[mi]
    pair_bytes = 2 * m.blk_stream_bytes
    b0 = m.byte_base + (gp - m.gp_base) * pair_bytes
    c_first = ((b0 + 256 * m.reg[0][1]) * C - 1) // B
    c_last = min(C - 1, ((b0 + pair_bytes) * C - 1) // B)
    return c_first, c_last


CASES = [
    ("4096x4096 [5,4]", [Mat(32, [(0, 5), (16, 4)], 128)], 128),
    ("qkv fused", [Mat(32, [(0, 5), (16, 4)], 128), Mat(32, [(0, 4)], 128), Mat(32, [(0, 6), (8, 5), (40, 4)], 128)], 128),
    ("gate|up 4096x11008 [4,3]", [Mat(86, [(0, 4), (16, 3)], 128)] * 1 + [Mat(86, [(0, 4), (16, 3)], 128)], 128),
    ("down 11008x4096", [Mat(32, [(0, 5), (36, 4)], 344)], 344),
    ("head 4096x32000 6-bit", [Mat(250, [(0, 6)], 128)], 128),
    ("tiny 256x128 8-bit", [Mat(1, [(0, 8)], 8)], 8),
    ("tiny 64x96", [Mat(1, [(0, 4)], 2)], 2),
    ("K=2048 vocab", [Mat(250, [(0, 6)], 64)], 64),
    ("2-bit 8192x8192", [Mat(64, [(0, 3), (80, 2)], 256)], 256),
]


@pytest.mark.parametrize("name,mats,KS", CASES, ids=[c[0] for c in CASES])
def test_work_split(name, mats, KS):
    B, GP, C, starts = split(mats, KS)
    assert starts[0] == 0 and all(b > a for a, b in zip(starts, starts[1:])), "every CTA owns at least one unit, in order"
    for gp in range(GP):
        lo, hi = gp * KS, (gp + 1) * KS
        touching = [c for c in range(C) if starts[c] < hi and starts[c + 1] > lo]
        c_first, c_last = owners(mats, KS, B, C, gp)
        assert touching == list(range(c_first, c_last + 1)), (name, gp)
        max_stream = max(m.blk_stream_bytes for m in mats)
        assert c_last - c_first + 1 <= (2 * max_stream * C) // B + 2
    # per-CTA row capacity bound used by the launcher
    max_bits = max(b for m in mats for _, b, _ in m.reg)
    min_bits = min(b for m in mats for _, b, _ in m.reg)
    per_cta = (B // C + 256 * max_bits + 256 * min_bits - 1) // (256 * min_bits) + 2
    assert max(b - a for a, b in zip(starts, starts[1:])) <= per_cta
