"""GPU cross-check against the UNMODIFIED reference extension (oracle/_ref/exllamav2_ext_ref.so, built from
/root/reference by oracle/build_ref.py; skipped when it is absent) on identical tensors -- SURVEY.md 8c row (iv).

Contract (SURVEY.md 8c tolerance row):
  reconstruct                bit-exact
  gemm (M <= 32, force_cuda) rel_l2(new, ref) <= 1e-3  and  rel_l2(new, truth) <= rel_l2(ref, truth) + 1e-5
  rms_norm                   <= 1 fp16 ulp;   rope: bit-exact;   Q4 kv pack/unpack: bit-exact (same intrinsics)
"""
import numpy as np
import pytest
import torch

import cases
import exl2_oracle as oracle

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


@pytest.fixture(scope="module")
def ref():
    from build_ref import load_ref
    m = load_ref()
    if m is None:
        pytest.skip("reference extension not built (oracle/_ref)")
    return m


def _mine(name):
    from exllamav2_b200.linear import ExLlamaV2Linear, load_tensor_dict
    w_np = cases.make_case(name)
    K, N = cases.case_shape(name)
    lin = ExLlamaV2Linear(K, N, has_bias="bias" in w_np, device=DEV)
    lin.load(load_tensor_dict(w_np, DEV))
    return lin, w_np


# the reference's kernels need N % 32 == 0 style tiles; keep to the shapes it supports
REF_CASES = [n for n in list(cases.EXL2_CASES) + list(cases.GPTQ_CASES) if cases.case_shape(n)[1] % 32 == 0]


@pytest.mark.parametrize("name", REF_CASES)
def test_reconstruct_matches_reference(ref, name):
    from gen_golden import ref_make_q_matrix
    lin, w_np = _mine(name)
    h, keep, temp_dq, (K, N) = ref_make_q_matrix(ref, w_np)
    W_ref = torch.empty((K, N), dtype=torch.half, device=DEV)
    ref.reconstruct(h, W_ref)
    W_new = lin.get_weight_tensor_dq()
    assert torch.equal(W_ref.view(torch.int16), W_new.view(torch.int16))
    ref.free_q_matrix(h)
    lin.unload()


@pytest.mark.parametrize("name", REF_CASES)
@pytest.mark.parametrize("M", [1, 4, 8, 19])
def test_gemm_matches_reference(ref, name, M):
    from gen_golden import ref_make_q_matrix
    lin, w_np = _mine(name)
    h, keep, temp_dq, (K, N) = ref_make_q_matrix(ref, w_np)
    a = cases.activations(name, M)
    at = torch.from_numpy(a).to(DEV)
    c_ref = torch.empty((M, N), dtype=torch.half, device=DEV)
    ref.gemm_half_q_half(at, h, c_ref, True)
    c_new = lin.forward(at)
    W = oracle.exl2_reconstruct(w_np) if name in cases.EXL2_CASES else oracle.gptq_reconstruct(w_np)
    truth = oracle.gemm_truth(a, W, w_np.get("bias"))
    e_ref = oracle.rel_l2(c_ref.cpu().numpy(), truth)
    e_new = oracle.rel_l2(c_new.cpu().numpy(), truth)
    e_x = oracle.rel_l2(c_new.cpu().numpy(), c_ref.float().cpu().numpy())
    assert e_x <= 1e-3 + e_ref, f"new vs ref {e_x:.2e} (ref vs truth {e_ref:.2e})"
    assert e_new <= e_ref + 1e-5, f"new {e_new:.2e} is further from truth than ref {e_ref:.2e}"
    ref.free_q_matrix(h)
    lin.unload()


def test_ops_match_reference(ref):
    from exllamav2_b200 import ext as ext_c
    from exllamav2_b200.ext import none_tensor
    rng = np.random.default_rng(123)
    # rms_norm
    x = torch.from_numpy(rng.normal(0, 1.5, size=(5, 4096)).astype(np.float16)).to(DEV)
    w = torch.from_numpy((1 + 0.1 * rng.normal(size=(4096,))).astype(np.float16)).to(DEV)
    y_ref, y_new = torch.empty_like(x), torch.empty_like(x)
    ref.rms_norm(x, w, y_ref, 1e-5)
    ext_c.rms_norm(x, w, y_new, 1e-5)
    diff = (y_ref.view(torch.int16).int() - y_new.view(torch.int16).int()).abs()
    assert diff.max().item() <= 1
    # rope (both styles), with per-batch offsets
    hd, heads = 128, 8
    sin, cos = oracle.rope_tables(hd, 128)
    st, ct = torch.from_numpy(sin).to(DEV), torch.from_numpy(cos).to(DEV)
    offs = torch.tensor([0, 11], dtype=torch.int, device=DEV)
    for neox in (True, False):
        xr = torch.from_numpy(rng.normal(0, 1, size=(2, 6, heads * hd)).astype(np.float16)).to(DEV)
        a, b = xr.clone(), xr.clone()
        ref.rope_(a, st, ct, 17, heads, hd, offs, neox)
        ext_c.rope_(b, st, ct, 17, heads, hd, offs, neox)
        assert torch.equal(a.view(torch.int16), b.view(torch.int16)), f"rope neox={neox}"
    # Q4 kv
    k = torch.from_numpy(rng.normal(0, 1, size=(2, 9, 8, 128)).astype(np.float16)).to(DEV)
    v = torch.from_numpy(rng.normal(0, 3, size=(2, 9, 8, 128)).astype(np.float16)).to(DEV)
    outs = []
    for e in (ref, ext_c):
        kq = torch.zeros((2, 9, 8, 64), dtype=torch.uint8, device=DEV)
        ks = torch.zeros((2, 9, 8, 4), dtype=torch.half, device=DEV)
        vq, vs = torch.zeros_like(kq), torch.zeros_like(ks)
        e.fp16_to_q_kv(k, kq, ks, v, vq, vs, 2, 2, 6, 0, none_tensor, none_tensor, 4)
        ko, vo = torch.zeros_like(k), torch.zeros_like(v)
        e.q_to_fp16_kv(kq, ko, ks, vq, vo, vs, 2, 2, 6, 0, none_tensor, none_tensor, 4)
        outs.append((kq, ks, vq, vs, ko, vo))
    for r, n in zip(outs[0], outs[1]):
        assert torch.equal(r.view(torch.uint8), n.view(torch.uint8))
