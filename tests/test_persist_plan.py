"""Work tables for the persistent decode kernel planned for round 2 (exllamav2_b200/persist_plan.py): every strip's K range
is tiled exactly once, boundaries are quantisation-group starts, contributor bookkeeping is consistent, every CTA has an
entry in every phase."""
import os
import random
import sys

sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "oracle"))

import synth
from exllamav2_b200.persist_plan import MatrixShape, choose_plan, first_requests, llama_layer_phases, plan_aligned, plan_phase, snap


def _check_phase(mats, table, G):
    assert len(table) == G
    cover = {}
    for cta, segs in enumerate(table):
        for s in segs:
            st = mats[s.mat].group_starts()
            assert s.ks0 in st and s.ks1 in st and s.ks0 <= s.ks1
            assert s.first_cta <= cta < s.first_cta + s.n_contrib
            cover.setdefault((s.mat, s.strip), []).append((s.ks0, s.ks1, cta))
    for mi, m in enumerate(mats):
        for strip in range(m.strips):
            parts = sorted(cover[(mi, strip)], key=lambda t: (t[2]))
            pos = 0
            for ks0, ks1, _ in parts:                      # in CTA order the ranges tile [0, KS) without gaps or overlap
                assert ks0 == pos or ks0 == ks1, (m.name, strip, parts)
                pos = max(pos, ks1)
            assert pos == m.KS
            ctas = [c for _, _, c in parts]
            seg0 = next(s for s in table[ctas[0]] if (s.mat, s.strip) == (mi, strip))
            assert ctas == list(range(seg0.first_cta, seg0.first_cta + seg0.n_contrib))     # every contributor shows up, once


def test_llama7b_layer_tables():
    gp = lambda name, K: synth.group_plan(K, [5, 4], [0.1, 0.9], 128)
    phases = llama_layer_phases(4096, 11008, 32, 32, 128, gp)
    G = 296
    for mats in phases:
        table = choose_plan(mats, G)
        _check_phase(mats, table, G)
        work = [sum(s.ks1 - s.ks0 for s in segs) for segs in table]
        busy = [w for w in work if w]
        assert max(busy) <= 1.35 * (sum(work) / len(busy)) + 8        # balanced up to group snapping
        req = first_requests(mats, table[0], 4)
        assert req and all(b > a for _, _, a, b in req) and len(req) <= 8


def test_random_shapes_and_group_mixes():
    rnd = random.Random(7)
    for _ in range(60):
        K = 32 * rnd.choice([4, 8, 16, 43, 64, 128])
        nm = rnd.choice([1, 2, 3])
        mats = []
        for i in range(nm):
            bits = rnd.choice([(4,), (5, 4), (4, 3), (8, 6, 5), (6, 3, 2)])
            prop = {1: (1.0,), 2: (0.1, 0.9), 3: (0.05, 0.15, 0.8)}[len(bits)]
            gs = rnd.choice([32, 64, 128])
            mats.append(MatrixShape(f"m{i}", K, 8 * rnd.randint(1, 200), tuple(synth.group_plan(K, list(bits), list(prop), gs))))
        G = rnd.choice([1, 7, 148, 296])
        for table in (plan_phase(mats, G), plan_aligned(mats, G)):
            if table is None:
                continue
            _check_phase(mats, table, len(table))
        _check_phase(mats, choose_plan(mats, G), G)


def test_snap():
    st = [0, 4, 8, 9, 10, 14, 16]
    assert [snap(st, k) for k in (0, 3, 4, 8, 9, 13, 15, 16, 20)] == [0, 0, 4, 8, 9, 10, 14, 16, 16]
