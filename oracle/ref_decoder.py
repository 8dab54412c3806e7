"""The REAL competitor, end to end: a Llama decode step on the UNMODIFIED reference extension (oracle/_ref), on the same
synthetic tensors, same GPU, same process as bench.py's own arm.

TEST / BENCH INFRASTRUCTURE -- never imported by the product path.  bench.py's `reference_cuda_ext` leg and
tools/ use it; it needs oracle/_ref/exllamav2_ext_ref.so (python oracle/build_ref.py).

Per layer it issues exactly the reference's own call sequence for a quantized Llama block with a Q4 cache
(exllamav2/attn.py:466-638 forward_paged, mlp.py:318-358, cache.py:472-556):
    q_to_fp16_kv (whole live cache -> fp16 temp)  ->  q_attn_forward_1 (rms_norm, q/k/v gemm, rope)
    ->  flash_attn_with_kvcache on the fp16 temp (third-party; torch SDPA when flash-attn does not run on this GPU)
    ->  fp16_to_q_kv (new rows)  ->  q_attn_forward_2  ->  q_mlp_forward_
then rms_norm + gemm_half_q_half for the head -- through the reference's fused QAttn / QMLP handles, so its autotuner
(q_gemm_autotune.cuh:11) and kernel selection are its own.  The host loop is replayed from ONE CUDA graph (more than the
stock Python loop gets: the reference captures per-module graphs only), so the number is the reference's kernels
without its Python overhead.
"""
from __future__ import annotations

import math
import os
import sys

import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
for p in (ROOT, HERE):
    if p not in sys.path:
        sys.path.insert(0, p)

PAGE = 256
none_tensor = torch.empty((1, 1), device="meta")


def _ref_q_matrix(ref, w: dict, K: int, dev):
    """exllamav2/ext.py:325-360 (EXL2 branch) driving the reference extension."""
    w = dict(w)
    w["q_scale_max"] = w["q_scale_max"] * (1.0 / 256)
    w["q_perm"] = w["q_perm"].short()
    w["q_invperm"] = w["q_invperm"].short()
    w["q_group_map"] = ref.make_group_map(w["q_groups"].cpu(), w["q_weight"].shape[0]).to(dev)
    h = ref.make_q_matrix(w["q_weight"], w["q_perm"], w["q_invperm"], w["q_scale"], w["q_scale_max"], w["q_groups"], w["q_group_map"],
                          none_tensor, none_tensor, none_tensor, none_tensor, none_tensor, 0)
    return h, w


class RefDecoder:
    def __init__(self, ref, cfg, device="cuda:0", seed: int = 0, cache_len: int = 1024):
        from exllamav2_b200 import synthetic
        from exllamav2_b200.model import rope_tables
        self.ref, self.cfg, self.dev = ref, cfg, torch.device(device)
        dev = self.dev
        H, KVH, hd, hid, inter = cfg.num_heads, cfg.num_kv_heads, cfg.head_dim, cfg.hidden_size, cfg.intermediate_size
        gen = torch.Generator(device=dev)
        gen.manual_seed(seed)
        self.keep, self.layers, self.handles = [], [], []

        def lin(K, N, plan, s, perm_seed=None):
            bits, prop, gs = plan
            w = synthetic.random_exl2(K, N, bits, prop, gs, device=dev, seed=s, weight_std=1.0 / math.sqrt(K), perm_seed=perm_seed)
            h, kw = _ref_q_matrix(ref, w, K, dev)
            self.keep.append(kw)
            self.handles.append(h)
            return h

        s = seed * 100003
        max_rows = 16
        self.temp_state = torch.empty((max_rows, max(hid, inter)), dtype=torch.half, device=dev)
        self.temp_a = torch.empty((max_rows, inter), dtype=torch.half, device=dev)
        self.temp_b = torch.empty((max_rows, inter), dtype=torch.half, device=dev)
        self.temp_dq = torch.empty((max(hid * inter, hid * cfg.vocab_size),), dtype=torch.half, device=dev)
        for li in range(cfg.num_layers):          # same seeds as exllamav2_b200.model.ExLlamaV2Decoder
            mp = cfg.plan.mlp[li % len(cfg.plan.mlp)]
            q, k = lin(hid, H * hd, cfg.plan.attn, s + 1, s + 1), lin(hid, KVH * hd, cfg.plan.attn, s + 2, s + 1)
            v, o = lin(hid, KVH * hd, cfg.plan.attn, s + 3, s + 1), lin(H * hd, hid, cfg.plan.attn, s + 4)
            g, u, d = lin(hid, inter, mp, s + 5, s + 5), lin(hid, inter, mp, s + 6, s + 5), lin(inter, hid, mp, s + 7)
            s += 16
            n1 = (1 + 0.1 * torch.randn((hid,), device=dev, generator=gen)).half()
            n2 = (1 + 0.1 * torch.randn((hid,), device=dev, generator=gen)).half()
            # make_q_attn / make_q_mlp as attn.py:300-330 / mlp.py:204-223 call them; use_graphs = False: the whole step is
            # captured in one CUDA graph below instead of the reference's per-module graphs
            attn = ref.make_q_attn(n1, none_tensor, True, False, cfg.norm_eps, q, k, v, o, self.temp_state, self.temp_dq, max_rows, hid, H, KVH,
                                   hd, cfg.max_seq_len, True, 2, hd, none_tensor, none_tensor, none_tensor, none_tensor, False, False)
            mlp = ref.make_q_mlp(n2, none_tensor, True, cfg.norm_eps, g, u, d, self.temp_state, self.temp_a, self.temp_b, self.temp_dq,
                                 max_rows, False, True, none_tensor, none_tensor, False, False)
            self.layers.append((attn, mlp, n1, n2))
        self.final_norm = (1 + 0.1 * torch.randn((hid,), device=dev, generator=gen)).half()
        self.lm_head = lin(hid, cfg.vocab_size, cfg.plan.head, s + 9)
        self.embed = (0.02 * torch.randn((cfg.vocab_size, hid), device=dev, generator=gen)).half()
        self.sin, self.cos = rope_tables(hd, cfg.max_seq_len, cfg.rope_theta, dev)
        pages = cache_len // PAGE
        shp = (pages, PAGE, KVH, hd)
        L = cfg.num_layers
        self.kq = [torch.zeros(shp[:3] + (hd // 2,), dtype=torch.uint8, device=dev) for _ in range(L)]
        self.vq = [torch.zeros_like(self.kq[0]) for _ in range(L)]
        self.ks = [torch.zeros(shp[:3] + (hd // 32,), dtype=torch.half, device=dev) for _ in range(L)]
        self.vs = [torch.zeros_like(self.ks[0]) for _ in range(L)]
        self.temp_k = torch.zeros(shp, dtype=torch.half, device=dev)
        self.temp_v = torch.zeros(shp, dtype=torch.half, device=dev)
        self.block_table = torch.arange(pages, dtype=torch.int32, device=dev).view(1, pages)
        self.seqlens = torch.zeros((1,), dtype=torch.int32, device=dev)
        self.ids = torch.zeros((1, 1), dtype=torch.long, device=dev)
        self.x = torch.empty((1, 1, hid), dtype=torch.half, device=dev)
        self.q = torch.empty((1, 1, H, hd), dtype=torch.half, device=dev)
        self.k = torch.empty((1, 1, KVH, hd), dtype=torch.half, device=dev)
        self.v = torch.empty((1, 1, KVH, hd), dtype=torch.half, device=dev)
        self.xn = torch.empty((1, hid), dtype=torch.half, device=dev)
        self.logits = torch.empty((1, cfg.vocab_size), dtype=torch.half, device=dev)
        self.pos_idx = torch.arange(cache_len, device=dev)
        self.attention = "flash_attn_with_kvcache"
        try:
            from flash_attn import flash_attn_with_kvcache
            self._fa = flash_attn_with_kvcache
            self._attn_fa(0)                                         # does it run on this GPU?
            torch.cuda.synchronize()
        except Exception as e:     # noqa: BLE001
            self._fa = None
            self.attention = f"torch SDPA over the fp16 temp (flash-attn unavailable: {str(e)[:80]})"
        self.seqlens.zero_()
        self.graph = None

    # -- attention on the fp16 temp cache, appending the new K/V row ----------------------------------------------------------
    def _attn_fa(self, li):
        hd = self.cfg.head_dim
        return self._fa(q=self.q, k=self.k, v=self.v, k_cache=self.temp_k, v_cache=self.temp_v, cache_seqlens=self.seqlens,
                        block_table=self.block_table, causal=True, softmax_scale=1.0 / math.sqrt(hd))

    def _attn_sdpa(self, li):
        cfg = self.cfg
        H, KVH, hd = cfg.num_heads, cfg.num_kv_heads, cfg.head_dim
        S = self.temp_k.shape[0] * PAGE
        kc, vc = self.temp_k.view(1, S, KVH, hd), self.temp_v.view(1, S, KVH, hd)
        idx = self.seqlens.long()
        kc[0].index_copy_(0, idx, self.k.view(1, KVH, hd))
        vc[0].index_copy_(0, idx, self.v.view(1, KVH, hd))
        mask = (self.pos_idx <= idx).view(1, 1, 1, S)
        rep = H // KVH
        kk = kc.transpose(1, 2).repeat_interleave(rep, dim=1) if rep > 1 else kc.transpose(1, 2)
        vv = vc.transpose(1, 2).repeat_interleave(rep, dim=1) if rep > 1 else vc.transpose(1, 2)
        o = torch.nn.functional.scaled_dot_product_attention(self.q.transpose(1, 2), kk, vv, attn_mask=mask)
        return o.transpose(1, 2)

    def step(self):
        ref, cfg = self.ref, self.cfg
        H, hd = cfg.num_heads, cfg.head_dim
        torch.index_select(self.embed, 0, self.ids.view(-1), out=self.x.view(1, -1))
        for li, (attn, mlp, n1, n2) in enumerate(self.layers):
            ref.q_to_fp16_kv(self.kq[li], self.temp_k, self.ks[li], self.vq[li], self.temp_v, self.vs[li], 1, 0, 0, PAGE,
                             self.seqlens, self.block_table, 4)
            ref.q_attn_forward_1(attn, self.x, 1, 1, 0, self.seqlens, self.q, self.k, self.v, self.sin, self.cos, [], none_tensor)
            ao = self._attn_fa(li) if self._fa is not None else self._attn_sdpa(li)
            ref.fp16_to_q_kv(self.temp_k, self.kq[li], self.ks[li], self.temp_v, self.vq[li], self.vs[li], 1, 0, 1, PAGE,
                             self.seqlens, self.block_table, 4)
            ref.q_attn_forward_2(attn, self.x, ao.reshape(1, 1, H * hd), 1, 1, [], none_tensor)
            ref.q_mlp_forward_(mlp, self.x.view(1, -1), [], none_tensor)
        self.seqlens.add_(1)
        ref.rms_norm(self.x.view(1, -1), self.final_norm, self.xn, cfg.norm_eps)
        ref.gemm_half_q_half(self.xn, self.lm_head, self.logits, False)

    def capture(self, argmax: bool = True):
        s = torch.cuda.Stream(self.dev)
        saved = self.seqlens.clone()

        def body():
            self.step()
            if argmax:
                torch.argmax(self.logits, dim=-1, keepdim=True, out=self.ids)
        with torch.cuda.stream(s):
            for _ in range(3):
                body()
            torch.cuda.synchronize()
            self.seqlens.copy_(saved)
            g = torch.cuda.CUDAGraph()
            with torch.cuda.graph(g, stream=s):
                body()
        torch.cuda.synchronize()
        self.seqlens.copy_(saved)
        self.graph = g

    def free(self):
        for attn, mlp, _, _ in self.layers:
            self.ref.free_q_attn(attn)
            self.ref.free_q_mlp(mlp)
        for h in self.handles:
            self.ref.free_q_matrix(h)
        self.layers, self.handles, self.keep = [], [], []


def time_reference_decode(cfg, prompt_len: int = 128, steps: int = 64, warmup: int = 8, device="cuda:0") -> dict:
    """tokens/s of the reference extension on the synthetic model: the stock-like eager loop and the whole-step graph."""
    from build_ref import load_ref
    ref = load_ref()
    if ref is None:
        return {"unavailable": "oracle/_ref/exllamav2_ext_ref.so not built"}
    dec = RefDecoder(ref, cfg, device=device, seed=0)
    g = torch.Generator(device="cpu").manual_seed(0)
    prompt = torch.randint(0, cfg.vocab_size, (1, prompt_len), generator=g).to(device)
    # feed the prompt one token at a time (> 210 calls per matrix shape: the autotuner settles, q_gemm_autotune.cuh)
    for t in range(prompt_len):
        dec.ids.copy_(prompt[:, t:t + 1])
        dec.step()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)

    def eager():
        dec.step()
        torch.argmax(dec.logits, dim=-1, keepdim=True, out=dec.ids)
    for _ in range(warmup):
        eager()
    torch.cuda.synchronize()
    e0.record()
    for _ in range(steps):
        eager()
    e1.record()
    torch.cuda.synchronize()
    ms_eager = e0.elapsed_time(e1) / steps
    out = {"eager_tokens_per_s": 1000.0 / ms_eager, "eager_ms_per_step": ms_eager, "attention": dec.attention,
           "sequence": "q_to_fp16_kv, q_attn_forward_1, attention on the fp16 temp, fp16_to_q_kv, q_attn_forward_2, q_mlp_forward_ per layer; rms_norm + gemm_half_q_half head"}
    try:
        dec.capture()
        for _ in range(warmup):
            dec.graph.replay()
        torch.cuda.synchronize()
        e0.record()
        for _ in range(steps):
            dec.graph.replay()
        e1.record()
        torch.cuda.synchronize()
        ms_graph = e0.elapsed_time(e1) / steps
        out.update(graph_tokens_per_s=1000.0 / ms_graph, graph_ms_per_step=ms_graph)
    except Exception as e:     # noqa: BLE001
        out["graph_error"] = str(e)[:200]
    out["finite"] = bool(torch.isfinite(dec.logits).all())
    dec.free()
    del dec
    torch.cuda.empty_cache()
    return out


if __name__ == "__main__":
    import json

    from exllamav2_b200.model import PRESETS
    print(json.dumps(time_reference_decode(PRESETS[sys.argv[1] if len(sys.argv) > 1 else "llama2-7b-4.0bpw"]())))
