"""Work tables for a persistent decode kernel (host-side logic only -- no kernel consumes them yet).

The per-launch kernel computes its unit range from blockIdx (gemm_tc.cu: units = (matrix, strip, slab), CTA i covers
[i*U/G, (i+1)*U/G), boundaries snapped to quantisation-group starts).  A persistent kernel runs every phase of a decoder
layer with the SAME G resident CTAs and must know, for every (phase, CTA), its segments up front -- to request the next
phase's weights before it waits at the phase barrier, and because every CTA has to arrive at every barrier even when it
has no work.  This module builds those tables and is exercised on the CPU (tests/test_persist_plan.py).

A phase is a list of matrices that share K (e.g. [q, k, v], [gate, up]); a matrix is described by its width N and its
quantisation groups along K as (bits, rows) pairs (synthetic.group_plan / the q_groups tensor).  Everything is in slabs of
32 rows and strips of 128 columns, like the kernel.
"""
from __future__ import annotations

from dataclasses import dataclass

SLAB_K = 32
STRIP_N = 128


@dataclass(frozen=True)
class MatrixShape:
    name: str
    K: int
    N: int
    groups: tuple            # ((bits, rows), ...) along K, rows multiples of 32

    @property
    def KS(self) -> int:
        return self.K // SLAB_K

    @property
    def strips(self) -> int:
        return (self.N + STRIP_N - 1) // STRIP_N

    def group_starts(self) -> list[int]:
        """first slab of every group, plus KS at the end"""
        out, ks = [], 0
        for _, rows in self.groups:
            assert rows % SLAB_K == 0, "group rows must be multiples of 32"
            out.append(ks)
            ks += rows // SLAB_K
        assert ks == self.KS, f"{self.name}: groups cover {ks} slabs, K/32 = {self.KS}"
        return out + [self.KS]

    def packed_bytes_of(self, ks0: int, ks1: int) -> int:
        """packed weight bytes of slabs [ks0, ks1) of ONE strip (128 columns)"""
        total, ks = 0, 0
        for bits, rows in self.groups:
            n = rows // SLAB_K
            lo, hi = max(ks, ks0), min(ks + n, ks1)
            if hi > lo:
                total += (hi - lo) * SLAB_K * STRIP_N * bits // 8
            ks += n
        return total


@dataclass(frozen=True)
class Segment:
    mat: int                 # index into the phase's matrix list
    strip: int
    ks0: int                 # slab range, group-aligned; may be empty (ks0 == ks1): the CTA still takes part in the
    ks1: int                 #   strip's split-K hand-off, exactly like the per-launch kernel
    first_cta: int           # the strip's contributors are CTAs first_cta .. first_cta + n_contrib - 1
    n_contrib: int


def snap(starts: list[int], ks: int) -> int:
    """first slab of the group containing slab ks (ks == KS maps to KS) -- tc_group_start in gemm_tc.cu"""
    lo, hi = 0, len(starts) - 1
    if ks >= starts[-1]:
        return starts[-1]
    while hi - lo > 1:
        mid = (lo + hi) // 2
        if starts[mid] <= ks:
            lo = mid
        else:
            hi = mid
    return starts[lo]


def cta_of_unit(x: int, G: int, U: int) -> int:
    return ((x + 1) * G - 1) // U


def plan_phase(mats: list[MatrixShape], G: int) -> list[list[Segment]]:
    """Stream-K over all units of the phase (the per-launch kernel's rule, so both agree bit for bit): table[cta] = segments."""
    assert mats and all(m.KS == mats[0].KS for m in mats), "matrices of a phase share K"
    KS = mats[0].KS
    unit_begin, U = [], 0
    for m in mats:
        unit_begin.append(U)
        U += m.strips * KS
    G = min(G, U)
    starts = [m.group_starts() for m in mats]
    table: list[list[Segment]] = []
    for cta in range(G):
        u0, u1 = cta * U // G, (cta + 1) * U // G
        segs, u = [], u0
        while u < u1:
            mi = max(i for i in range(len(mats)) if u >= unit_begin[i])
            local = u - unit_begin[mi]
            strip, ks_a = divmod(local, KS)
            seg = min(KS - ks_a, u1 - u)
            sb = unit_begin[mi] + strip * KS
            first, last = cta_of_unit(sb, G, U), cta_of_unit(sb + KS - 1, G, U)
            segs.append(Segment(mi, strip, snap(starts[mi], ks_a), snap(starts[mi], ks_a + seg), first, last - first + 1))
            u += seg
        table.append(segs)
    return table


def plan_aligned(mats: list[MatrixShape], G: int) -> list[list[Segment]] | None:
    """Every strip cut into S K-ranges, one segment per CTA (S = G // strips, need not divide KS).  None if G < strips."""
    KS = mats[0].KS
    strips = sum(m.strips for m in mats)
    if strips > G:
        return None
    S = min(G // strips, max(1, KS // 8))
    return plan_phase(mats, strips * S)


def choose_plan(mats: list[MatrixShape], G: int, fixed_cost_slabs: int = 16) -> list[list[Segment]]:
    """The launcher's cost model (gemm_tc_launch): slowest CTA = segments * F + slabs; pad the table to G CTAs with
    empty entries (a persistent CTA without work in a phase still arrives at the phase barrier)."""
    stream = plan_phase(mats, G)
    cost = lambda t: max((len(s) * fixed_cost_slabs + sum(x.ks1 - x.ks0 for x in s)) for s in t)
    best = stream
    al = plan_aligned(mats, G)
    if al is not None and cost(al) <= cost(stream):
        best = al
    return best + [[] for _ in range(G - len(best))]


def first_requests(mats: list[MatrixShape], segs: list[Segment], stages: int) -> list[tuple[int, int, int, int]]:
    """What a CTA asks for BEFORE waiting at the barrier in front of this phase: (mat, strip, ks0, ks1) of the first `stages`
    groups of each warpgroup's half of its first segment (weights never depend on the previous phase)."""
    if not segs:
        return []
    s = segs[0]
    st = mats[s.mat].group_starts()
    mid = snap(st, s.ks0 + ((s.ks1 - s.ks0 + 1) >> 1))
    out = []
    for lo, hi in ((s.ks0, mid), (mid, s.ks1)):
        ks, n = lo, 0
        while ks < hi and n < stages:
            nxt = st[st.index(ks) + 1]
            out.append((s.mat, s.strip, ks, nxt))
            ks, n = nxt, n + 1
    return out


def llama_layer_phases(hidden: int, inter: int, heads: int, kv_heads: int, head_dim: int, groups_of) -> list[list[MatrixShape]]:
    """The four dequant-GEMM phases of a decoder layer; groups_of(name, K) -> ((bits, rows), ...)."""
    mk = lambda name, K, N: MatrixShape(name, K, N, tuple(groups_of(name, K)))
    return [
        [mk("q", hidden, heads * head_dim), mk("k", hidden, kv_heads * head_dim), mk("v", hidden, kv_heads * head_dim)],
        [mk("o", heads * head_dim, hidden)],
        [mk("gate", hidden, inter), mk("up", hidden, inter)],
        [mk("down", inter, hidden)],
    ]
