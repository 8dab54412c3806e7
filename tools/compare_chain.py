"""Lock-step comparison of the chained and un-chained layer loops on test-tiny, stage by stage (diagnostics)."""
import math, sys
import numpy as np, torch
sys.path.insert(0, ".")
from exllamav2_b200 import ext as ext_c
from exllamav2_b200.model import ExLlamaV2Decoder, PRESETS

dev = "cuda:0"
def mk():
    d = ExLlamaV2Decoder(PRESETS["test-tiny"](), device=dev, seed=3, batch_size=1, cache_len=512)
    return d
A, B = mk(), mk()     # A: unchained fused-attn, B: chained
q_len = int(sys.argv[1]) if len(sys.argv) > 1 else 3
g = torch.Generator().manual_seed(1)
ids = torch.randint(0, 512, (1, q_len), generator=g).to(dev)
def rel(a, b):
    a, b = a.float(), b.float()
    return float((a - b).norm() / (b.norm() + 1e-9))
cfg = A.cfg
H, KVH, hd = cfg.num_heads, cfg.num_kv_heads, cfg.head_dim
def bufs(D):
    x = D.embed[ids].contiguous()
    q = torch.zeros((1, q_len, H * hd), dtype=torch.half, device=dev)
    k = torch.zeros((1, q_len, KVH * hd), dtype=torch.half, device=dev)
    return x, q, k, torch.zeros_like(k), torch.zeros_like(q)
xa, qa, ka, va, aoa = bufs(A)
xb, qb, kb, vb, aob = bufs(B)
n = len(A.layers)
for li in range(n):
    LA, LB = A.layers[li], B.layers[li]
    ext_c.q_attn_forward_1(LA.attn, xa, 1, q_len, -1, A.cache.cache_seqlens, qa, ka, va, A.sin, A.cos)
    ext_c.q_attn_forward_1_ex(LB.attn, xb, 1, q_len, -1, B.cache.cache_seqlens, qb, kb, vb, B.sin, B.cos, li > 0)
    torch.cuda.synchronize()
    print(f"L{li} qkv  ", rel(qb, qa), rel(kb, ka), rel(vb, va), bool(torch.isfinite(qb).all()))
    for D, q, k, v, ao, L, oc in ((A, qa, ka, va, aoa, LA, 0), (B, qb, kb, vb, aob, LB, LB.o_proj.q_handle)):
        c = D.cache
        ext_c.paged_attn_decode_q4(q.view(1, q_len, H, hd), k.view(1, q_len, KVH, hd), v.view(1, q_len, KVH, hd),
                                   c.key_states[li], c.key_scales[li], c.value_states[li], c.value_scales[li],
                                   c.cache_seqlens, c.block_table, ao.view(1, q_len, H, hd), 1.0 / math.sqrt(hd), oc)
    torch.cuda.synchronize()
    print(f"L{li} attn ", rel(aob, aoa))
    ext_c.q_attn_forward_2(LA.attn, xa, aoa, 1, q_len)
    ext_c.q_attn_forward_2_ex(LB.attn, xb, aob, 1, q_len, True, LB.chain_mlp)
    torch.cuda.synchronize()
    print(f"L{li} o    ", rel(xb, xa))
    ext_c.q_mlp_forward_(LA.mlp, xa)
    nxt = B.layers[li + 1].chain_attn if li + 1 < n else B.chain_head
    ext_c.q_mlp_forward_ex(LB.mlp, xb, True, nxt)
    torch.cuda.synchronize()
    print(f"L{li} mlp  ", rel(xb, xa), "temp_a", rel(LB.temp_a[:q_len], LA.temp_a[:q_len]))
if q_len == 1:
    ext_c.rms_norm(xa.view(1, -1), A.final_norm, A.xn, cfg.norm_eps)
    ext_c.gemm_half_q_half(A.xn, A.lm_head.q_handle, A.logits, False)
    ext_c.gemm_half_q_half_prepared(B.lm_head.q_handle, B.logits, True, cfg.norm_eps)
    torch.cuda.synchronize()
    print("head ", rel(B.logits, A.logits))
