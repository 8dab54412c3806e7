"""Host-side mirror of the reference's quantized-linear operator surface.

`make_q_matrix` follows exllamav2/ext.py:325-410 (tensor-dict -> handle; prescale folding, perm dtype, GPTQ v2 zero
offset, act-order detection) and `ExLlamaV2Linear` follows the quantized branch of exllamav2/linear.py
(load :119-171, forward :310-393, get_weight_tensor_dq :486-504, unload :230-239, tp_split :540-619).  Same names,
argument meaning and error behaviour, so tests read like the reference's tests/test_gemv.py.
"""
from __future__ import annotations

import torch

from . import ext as ext_c
from .ext import none_tensor


def make_q_matrix(w: dict, temp_dq: torch.Tensor = none_tensor, key: str | None = None, prescale: float = 1,
                  max_dq_rows: int = 0, offset_qzeros: bool = False) -> int:
    # EXL2
    if "q_weight" in w:
        w["q_scale_max"] *= prescale / 256                                  # ext.py:336
        if "q_perm" in w: w["q_perm"] = w["q_perm"].short()
        if "q_invperm" in w: w["q_invperm"] = w["q_invperm"].short()
        if "q_group_map" not in w:
            w["q_group_map"] = ext_c.make_group_map(w["q_groups"], w["q_weight"].shape[0]).to(w["q_groups"].device)
        return ext_c.make_q_matrix(
            w["q_weight"], w.get("q_perm", none_tensor), w.get("q_invperm", none_tensor), w["q_scale"],
            w["q_scale_max"], w["q_groups"], w["q_group_map"], none_tensor, none_tensor, none_tensor,
            w.get("bias", none_tensor), temp_dq, max_dq_rows)
    # GPTQ
    elif "qweight" in w:
        if prescale != 1: w["scales"] *= prescale
        if w["scales"].dtype == torch.float: w["scales"] = w["scales"].half()
        if offset_qzeros:
            w["qzeros"] -= 0b00010001000100010001000100010001                # ext.py:366-367
        if "g_idx" in w and not (w["g_idx"] == 0).all().item():             # act-order, ext.py:371
            w["q_perm"] = torch.empty((w["qweight"].shape[0] * 8,), dtype=torch.short, device=w["qweight"].device)
            w["q_invperm"] = torch.empty_like(w["q_perm"])
            return ext_c.make_q_matrix(
                w["qweight"], w["q_perm"], w["q_invperm"], none_tensor, none_tensor, none_tensor, none_tensor,
                w["qzeros"], w["scales"], w["g_idx"].cpu(), w.get("bias", none_tensor), temp_dq, max_dq_rows)
        return ext_c.make_q_matrix(
            w["qweight"], none_tensor, none_tensor, none_tensor, none_tensor, none_tensor, none_tensor,
            w["qzeros"], w["scales"], none_tensor, w.get("bias", none_tensor), temp_dq, max_dq_rows)
    raise ValueError("tensor dict is neither EXL2 nor GPTQ")


def load_tensor_dict(w_np: dict, device) -> dict:
    """numpy checkpoint tensors -> torch tensors on `device`, with the derived q_perm of module.py:118-121."""
    import numpy as np
    w = {}
    for k, v in w_np.items():
        t = torch.from_numpy(np.ascontiguousarray(v))
        w[k] = t if k == "g_idx" else t.to(device)
    if "q_invperm" in w:
        w["q_perm"] = torch.argsort(w["q_invperm"]).to(torch.int)           # module.py:120
    return w


def tp_column_slice(w_np: dict, a: int, b: int) -> dict:
    """Column shard [a, b) of a checkpoint tensor dict, cut BEFORE make_q_matrix (one process per GPU loads only
    its shard).  Mirrors what ExLlamaV2Linear.tp_split slices (linear.py:567-587: q_weight[:, a:b],
    q_scale[:, a/8:b/8], bias[a:b], shared q_scale_max/q_groups/q_invperm) and defines the GPTQ case the
    reference rejects (ext_qmatrix.cpp:130-135): qweight[:, a:b], qzeros[:, a/8:b/8], scales[:, a:b]."""
    assert a % 8 == 0 and b % 8 == 0
    out = dict(w_np)
    if "q_weight" in w_np:
        out["q_weight"] = w_np["q_weight"][:, a:b].copy()
        out["q_scale"] = w_np["q_scale"][:, a // 8:b // 8].copy()
    else:
        out["qweight"] = w_np["qweight"][:, a:b].copy()
        out["qzeros"] = w_np["qzeros"][:, a // 8:b // 8].copy()
        out["scales"] = w_np["scales"][:, a:b].copy()
    if "bias" in w_np:
        out["bias"] = w_np["bias"][a:b].copy()
    return out


class ExLlamaV2Linear:
    """Quantized linear layer (the q_handle branch of exllamav2/linear.py)."""

    def __init__(self, in_features: int, out_features: int, has_bias: bool = False, key: str = "linear",
                 prescale: float = 1, device="cuda:0"):
        self.in_features = in_features
        self.out_features = out_features
        self.has_bias = has_bias
        self.key = key
        self.prescale = prescale
        self._device = torch.device(device)
        self.q_handle = None
        self.q_tensors = None

    def device(self):
        return self._device

    def load(self, w: dict, offset_qzeros: bool = False):
        if self.has_bias:
            assert "bias" in w, self.key + " has no bias but bias expected"
        else:
            assert "bias" not in w, self.key + " has bias but bias is not expected"
        self.q_tensors = w                                                   # keeps the tensors alive, linear.py:145
        self.q_handle = make_q_matrix(w, none_tensor, prescale=self.prescale, offset_qzeros=offset_qzeros)
        self.prescale = 1

    def unload(self):
        if self.q_handle is not None:
            ext_c.free_q_matrix(self.q_handle)
            self.q_handle = None
        self.q_tensors = None

    def numel(self) -> int:
        return self.in_features * self.out_features

    def get_weight_tensor_dq(self) -> torch.Tensor:
        if self.q_handle is None:
            raise ValueError(f"Layer {self.key} has no data")
        tensor = torch.empty((self.in_features, self.out_features), dtype=torch.half, device=self.device())
        ext_c.reconstruct(self.q_handle, tensor)
        return tensor

    def forward(self, hidden_states: torch.Tensor, force_recons: bool = False, force_cuda: bool = False) -> torch.Tensor:
        if self.q_handle is None:
            raise ValueError(f"Layer {self.key} has no data")
        if force_recons:                                                     # linear.py:370-379
            return torch.matmul(hidden_states, self.get_weight_tensor_dq())
        output_shape = hidden_states.shape[:-1] + (self.out_features,)
        hs = hidden_states.view(-1, hidden_states.shape[-1])
        output = torch.empty((hs.shape[0], self.out_features), dtype=torch.half, device=self.device())
        ext_c.gemm_half_q_half(hs, self.q_handle, output, force_cuda)
        return output.view(output_shape)
