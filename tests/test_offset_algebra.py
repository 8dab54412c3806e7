"""The two-offset 4-bit unpack of gemm_tc_kernel (csrc/gemm_tc.cu, tc_dequant4): the tensor core multiplies the activations
with offset + q (1024 + q on even pair slots, 64 + q on odd ones) and the offsets and the zero point are removed afterwards
with two column-independent sums,   sum_k a_k (q_k - z) = D - (S1 + z * S0),   S1 = sum a_k * offset_k,  S0 = sum a_k.
Checked here in numpy with the tensor core's arithmetic (exact fp16 x fp16 products, fp32 accumulation): the identity holds
and the cancellation costs ~1e-5 of the result's scale, far below the fp16 rounding of the output (4.9e-4)."""
import numpy as np
import pytest


def _offsets(n_k):
    pair = (np.arange(n_k) // 2) % 16            # pair slot inside a 32-row slab
    return np.where(pair % 2 == 0, 1024.0, 64.0).astype(np.float32)


@pytest.mark.parametrize("z", [8, 1, 16])        # EXL2 zero point 8; GPTQ z + 1 in 1..16
@pytest.mark.parametrize("seed", [0, 1, 2])
def test_two_offset_identity(z, seed):
    rng = np.random.default_rng(seed)
    K, N = 128, 64                               # one quantisation group, 64 weight columns
    a = rng.normal(0, 1, size=(K,)).astype(np.float16)
    q = rng.integers(0, 16, size=(K, N))
    off = _offsets(K)
    A = (off[:, None] + q).astype(np.float16)    # what the unpack stores: exactly representable (<= 1039)
    assert np.array_equal(A.astype(np.float32), off[:, None] + q)
    prod = a.astype(np.float32)[:, None] * A.astype(np.float32)          # exact in fp32 (11 x 11 significant bits)
    D = np.zeros((N,), dtype=np.float32)
    for k in range(K):                           # fp32 accumulation, in order
        D += prod[k]
    a32 = a.astype(np.float32)
    S1 = np.float32(0)
    S0 = np.float32(0)
    for k in range(K):
        S1 = np.float32(S1 + a32[k] * off[k])
        S0 = np.float32(S0 + a32[k])
    got = D - (S1 + np.float32(z) * S0)
    want = (a.astype(np.float64)[:, None] * (q - z)).sum(0)
    scale = np.abs(a.astype(np.float64)).sum() * 8.0                    # what sum |a| |q - z| can reach
    assert np.max(np.abs(got - want)) < 2e-5 * scale * 128              # fp32 epsilon times the 1024-offset partial sums
    assert np.max(np.abs(got - want)) < 0.05 * np.std(want)             # and negligible against the result itself
