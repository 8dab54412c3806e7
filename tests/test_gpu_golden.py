"""Our CUDA path against the committed golden vectors of the reference extension (tests/golden)."""
import glob
import os

import numpy as np
import pytest
import torch

import cases
import exl2_oracle as oracle

pytestmark = pytest.mark.gpu
GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
LINEAR = sorted(os.path.basename(p)[len("linear_"):-4] for p in glob.glob(os.path.join(GOLD, "linear_*.npz")))
DEV = "cuda:0"


@pytest.mark.parametrize("name", LINEAR)
def test_linear_vs_golden(name):
    from exllamav2_b200.linear import ExLlamaV2Linear, load_tensor_dict
    g = np.load(os.path.join(GOLD, f"linear_{name}.npz"))
    w_np = cases.make_case(name)
    K, N = cases.case_shape(name)
    lin = ExLlamaV2Linear(K, N, has_bias="bias" in w_np, device=DEV)
    lin.load(load_tensor_dict(w_np, DEV))
    assert np.array_equal(lin.get_weight_tensor_dq().cpu().numpy().view(np.uint16), g["reconstruct"])
    for M in cases.M_VALUES:
        y = lin.forward(torch.from_numpy(cases.activations(name, M)).to(DEV)).cpu().numpy()
        ref = g[f"gemm_m{M}"].view(np.float16)
        assert oracle.rel_l2(y, ref.astype(np.float32)) <= 3e-3       # bounded by the reference's own fp16 accumulation error
    lin.unload()


def test_ops_vs_golden():
    from exllamav2_b200 import ext as ext_c
    from exllamav2_b200.ext import none_tensor
    g = np.load(os.path.join(GOLD, "ops.npz"))
    t = lambda k: torch.from_numpy(g[k].view(np.float16).copy()).to(DEV)
    y = torch.empty_like(t("norm_x"))
    ext_c.rms_norm(t("norm_x"), t("norm_w"), y, 1e-5)
    d = (y.view(torch.int16).int() - t("norm_y").view(torch.int16).int()).abs().max().item()
    assert d <= 1
    offs = torch.tensor([0, 5], dtype=torch.int, device=DEV)
    for tag, neox in (("neox", True), ("gptj", False)):
        x = t(f"rope_{tag}_x")
        ext_c.rope_(x, t("rope_sin"), t("rope_cos"), 9, 4, 128, offs, neox)
        assert np.array_equal(x.cpu().numpy().view(np.uint16), g[f"rope_{tag}_y"])
    k = t("kv_x")
    kq = torch.zeros((2, 4, 8, 64), dtype=torch.uint8, device=DEV)
    ks = torch.zeros((2, 4, 8, 4), dtype=torch.half, device=DEV)
    ext_c.fp16_to_q_kv(k, kq, ks, none_tensor, none_tensor, none_tensor, 2, 0, 4, 0, none_tensor, none_tensor, 4)
    assert np.array_equal(kq.cpu().numpy(), g["kv_q"]) and np.array_equal(ks.cpu().numpy().view(np.uint16), g["kv_s"])
    ko = torch.zeros_like(k)
    ext_c.q_to_fp16_kv(kq, ko, ks, none_tensor, none_tensor, none_tensor, 2, 0, 4, 0, none_tensor, none_tensor, 4)
    assert np.array_equal(ko.cpu().numpy().view(np.uint16), g["kv_y"])
