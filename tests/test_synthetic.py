"""The synthetic checkpoints bench.py and the decoder use (no network, no real weights): format-correct for the oracle,
and -- with weight_std -- dequantised weights of the requested spread, so that a random-init network keeps O(1)
activations through 32 layers in fp16 (an earlier version overflowed)."""
import math
import os
import sys

import numpy as np
import pytest
import torch

sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "oracle"))
import exl2_oracle as oracle
from exllamav2_b200 import synthetic


@pytest.mark.parametrize("bits,prop,gs", [((4,), (1.0,), 128), ((5, 4), (0.1, 0.9), 128), ((4, 3), (0.1, 0.9), 64), ((6,), (1.0,), 128)])
def test_random_exl2_weight_std(bits, prop, gs):
    K, N = 1024, 256
    target = 1.0 / math.sqrt(K)
    w = synthetic.random_exl2(K, N, bits, prop, gs, device="cpu", seed=5, weight_std=target)
    wn = {k: v.numpy() for k, v in w.items()}
    wn["q_groups"] = wn["q_groups"].astype(np.int16)
    W = oracle.exl2_reconstruct(wn).astype(np.float64)
    assert W.shape == (K, N) and np.isfinite(W).all()
    # weight_std is the nominal knob: the generator sizes q_scale_max from E[(s+1)^2] = 93.5, while the spread of the product
    # follows sqrt(E[(s+1)^4]) = 123.5, so the realised std is ~1.3-1.45x the nominal value (same for every matrix)
    assert 0.9 * target < W.std() < 1.6 * target, (W.std(), target)
    assert abs(W.mean()) < 0.3 * target                         # q - 2^(b-1) spans [-2^(b-1), 2^(b-1) - 1]: mean -0.5 steps
    # byte accounting used for the roofline: packed rows * N * 4 + scale words + scale_max + perm + one activation row + one output row
    b = synthetic.algorithmic_bytes(w, 1)
    assert b == w["q_weight"].numel() * 4 + w["q_scale"].numel() * 4 + w["q_scale_max"].numel() * 2 + 2 * K + 2 * K + 2 * N
    bpw = w["q_weight"].numel() * 32 / (K * N)
    assert min(bits) <= bpw <= max(bits)


def test_random_gptq_format():
    w = synthetic.random_gptq(512, 128, 128, device="cpu", seed=3, act_order=True)
    wn = {k: (v.numpy() if hasattr(v, "numpy") else v) for k, v in w.items()}
    W = oracle.gptq_reconstruct(wn)
    assert W.shape == (512, 128) and np.isfinite(W.astype(np.float64)).all()
