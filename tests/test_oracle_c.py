"""The C restatement (oracle/exl2_cpu.c, the CPU-baseline port) against the numpy oracle."""
import numpy as np
import pytest

import cases
import exl2_oracle as oracle


@pytest.fixture(scope="module")
def clib():
    import oracle_c
    return oracle_c.load()


@pytest.mark.parametrize("name", list(cases.EXL2_CASES))
def test_exl2_cpu_gemv(clib, name):
    import oracle_c
    w = cases.make_case(name)
    a = cases.activations(name, 1)
    y = oracle_c.exl2_gemv(clib, w, a[0])
    truth = oracle.gemm_truth(a, oracle.exl2_reconstruct(w))[0]
    assert oracle.rel_l2(y, truth) <= 5e-4      # weights are not rounded to fp16 per element here (fp32 scale product)


@pytest.mark.parametrize("name", ["gptq_g128", "gptq_g32"])
def test_gptq_cpu_gemv(clib, name):
    import oracle_c
    w = cases.make_case(name)
    a = cases.activations(name, 1)
    y = oracle_c.gptq_gemv(clib, w, a[0])
    truth = oracle.gemm_truth(a, oracle.gptq_reconstruct(w))[0]
    assert oracle.rel_l2(y, truth) <= 5e-4
