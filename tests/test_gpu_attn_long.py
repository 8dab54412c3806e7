"""Fused Q4 attention beyond short decode: split-KV over long contexts, fused RoPE, and the page-table guard.

  * contexts 1k / 4k / 16k (the cache is long enough that the launch uses several CTAs per head, csrc/attn_q4.cu split-KV):
    output vs fp64 attention over the oracle-dequantised cache, vectorised per head; also a SHORT context in the same long
    cache (one active chunk) and a context that is not a multiple of the chunk size;
  * fused RoPE: un-rotated q / k_new + tables == rope_ (bit-exact kernel, test_gpu_ops.py) followed by the plain call,
    bit for bit, including the cache bytes written;
  * a sequence that would run past its page table is refused (sticky status bit) and writes nothing.
"""
import numpy as np
import pytest
import torch

import exl2_oracle as oracle

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


def _setup(H, KVH, hd, pps, B, seed):
    page = 256
    rng = np.random.default_rng(seed)
    pages_total = B * pps
    block_table = rng.permutation(pages_total).reshape(B, pps).astype(np.int32)
    past_k = rng.normal(0, 1, size=(pages_total, page, KVH, hd)).astype(np.float16)
    past_v = rng.normal(0, 1, size=(pages_total, page, KVH, hd)).astype(np.float16)
    kq0, ks0 = oracle.kv_pack_q4(past_k)
    vq0, vs0 = oracle.kv_pack_q4(past_v)
    return page, rng, block_table, kq0, ks0, vq0, vs0


@pytest.mark.parametrize("H,KVH,hd,ctx,cache_len", [
    (32, 32, 128, 1000, 1024 * 2),      # 2 chunks
    (32, 8, 128, 4095, 4096),           # GQA, 8 chunks, ragged last chunk
    (8, 8, 64, 16000, 16384),           # 16k
    (32, 32, 128, 130, 16384),          # short context in a long cache: one active chunk, no merge
    (32, 32, 128, 513, 4096),           # just over one chunk
])
def test_split_kv_long_context(H, KVH, hd, ctx, cache_len):
    from exllamav2_b200 import ext as ext_c
    pps = cache_len // 256
    page, rng, block_table, kq0, ks0, vq0, vs0 = _setup(H, KVH, hd, pps, 1, ctx)
    q = rng.normal(0, 1, size=(1, 1, H, hd)).astype(np.float16)
    kn = rng.normal(0, 1, size=(1, 1, KVH, hd)).astype(np.float16)
    vn = rng.normal(0, 1, size=(1, 1, KVH, hd)).astype(np.float16)
    t = lambda a: torch.from_numpy(a).to(DEV)
    kq, ks, vq, vs = t(kq0.copy()), t(ks0.copy()), t(vq0.copy()), t(vs0.copy())
    out = torch.zeros((1, 1, H, hd), dtype=torch.half, device=DEV)
    sl = t(np.array([ctx], dtype=np.int32))
    for rep in range(2):          # twice: the merge counters must be back at zero after a launch
        kq.copy_(t(kq0)); ks.copy_(t(ks0)); vq.copy_(t(vq0)); vs.copy_(t(vs0))
        ext_c.paged_attn_decode_q4(t(q), t(kn), t(vn), kq, ks, vq, vs, sl, t(block_table), out, 1.0 / np.sqrt(hd))
        torch.cuda.synchronize()
        assert ext_c.paged_attn_status(DEV) == 0
        # gather this sequence's cached rows in position order, dequantise with the oracle
        pg = block_table[0, np.arange(ctx) // page]
        r = np.arange(ctx) % page
        kd = oracle.kv_unpack_q4(kq0[pg, r], ks0[pg, r]).astype(np.float64)       # [ctx, KVH, hd]
        vd = oracle.kv_unpack_q4(vq0[pg, r], vs0[pg, r]).astype(np.float64)
        K = np.concatenate([kd, kn[0].astype(np.float64)], axis=0)
        V = np.concatenate([vd, vn[0].astype(np.float64)], axis=0)
        group = H // KVH
        got = out[0, 0].cpu().numpy().astype(np.float64)
        for h in range(H):
            s = K[:, h // group] @ q[0, 0, h].astype(np.float64) / np.sqrt(hd)
            pr = np.exp(s - s.max())
            ref = (pr / pr.sum()) @ V[:, h // group]
            err = oracle.rel_l2(got[h], ref)
            assert err < 2e-3, (rep, h, err)
    # the appended row landed at position ctx
    nkq, nks = oracle.kv_pack_q4(kn)
    pgn = block_table[0, ctx // page]
    assert np.array_equal(kq.cpu().numpy()[pgn, ctx % page], nkq[0, 0])


@pytest.mark.parametrize("H,KVH,hd,neox", [(32, 32, 128, True), (32, 4, 64, True), (8, 8, 128, False), (8, 2, 64, False)])
def test_fused_rope_equals_rope_then_attention(H, KVH, hd, neox):
    from exllamav2_b200 import ext as ext_c
    page, rng, block_table, kq0, ks0, vq0, vs0 = _setup(H, KVH, hd, 2, 2, 7)
    seqlens = np.array([77, 300], dtype=np.int32)
    B = 2
    sin_np, cos_np = oracle.rope_tables(hd, 512)
    sin, cos = torch.from_numpy(sin_np).to(DEV), torch.from_numpy(cos_np).to(DEV)
    q = torch.from_numpy(rng.normal(0, 1, size=(B, 1, H, hd)).astype(np.float16)).to(DEV)
    kn = torch.from_numpy(rng.normal(0, 1, size=(B, 1, KVH, hd)).astype(np.float16)).to(DEV)
    vn = torch.from_numpy(rng.normal(0, 1, size=(B, 1, KVH, hd)).astype(np.float16)).to(DEV)
    t = lambda a: torch.from_numpy(a).to(DEV)
    sl, bt = t(seqlens), t(block_table)

    def run(fused):
        kq, ks, vq, vs = t(kq0.copy()), t(ks0.copy()), t(vq0.copy()), t(vs0.copy())
        out = torch.zeros((B, 1, H, hd), dtype=torch.half, device=DEV)
        if fused:
            ext_c.paged_attn_decode_q4(q, kn, vn, kq, ks, vq, vs, sl, bt, out, 1.0 / np.sqrt(hd), rope=(sin, cos, 2 if neox else 1))
        else:
            qr, kr = q.clone().view(B, 1, H * hd), kn.clone().view(B, 1, KVH * hd)
            ext_c.rope_(qr, sin, cos, -1, H, hd, sl, neox)          # past_len = -1: positions = cache_seqlens (rope.cu:39-43)
            ext_c.rope_(kr, sin, cos, -1, KVH, hd, sl, neox)
            ext_c.paged_attn_decode_q4(qr.view(B, 1, H, hd), kr.view(B, 1, KVH, hd), vn, kq, ks, vq, vs, sl, bt, out, 1.0 / np.sqrt(hd))
        torch.cuda.synchronize()
        return out, kq, ks
    o1, kq1, ks1 = run(True)
    o2, kq2, ks2 = run(False)
    assert torch.equal(kq1, kq2) and torch.equal(ks1.view(torch.int16), ks2.view(torch.int16))
    assert torch.equal(o1.view(torch.int16), o2.view(torch.int16))


def test_page_table_overrun_is_refused():
    from exllamav2_b200 import ext as ext_c
    H = KVH = 4
    hd = 64
    page, rng, block_table, kq0, ks0, vq0, vs0 = _setup(H, KVH, hd, 1, 1, 3)      # one page: 256 positions
    t = lambda a: torch.from_numpy(a).to(DEV)
    q = t(rng.normal(0, 1, size=(1, 2, H, hd)).astype(np.float16))
    kn = t(rng.normal(0, 1, size=(1, 2, KVH, hd)).astype(np.float16))
    vn = t(rng.normal(0, 1, size=(1, 2, KVH, hd)).astype(np.float16))
    kq, ks, vq, vs = t(kq0.copy()), t(ks0.copy()), t(vq0.copy()), t(vs0.copy())
    out = torch.zeros((1, 2, H, hd), dtype=torch.half, device=DEV)
    assert ext_c.paged_attn_status(DEV) == 0
    ext_c.paged_attn_decode_q4(q, kn, vn, kq, ks, vq, vs, t(np.array([255], dtype=np.int32)), t(block_table), out, 0.125)
    torch.cuda.synchronize()
    assert ext_c.paged_attn_status(DEV) & 1
    assert torch.equal(kq, t(kq0)) and torch.count_nonzero(out).item() == 0
    ext_c.paged_attn_clear_status(DEV)
    assert ext_c.paged_attn_status(DEV) == 0
