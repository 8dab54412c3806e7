"""ctypes binding of oracle/exl2_cpu.c (TEST INFRASTRUCTURE: tests/ and bench.py's CPU-baseline legs only)."""
import ctypes
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = os.path.join(HERE, "libexl2_cpu.so")


def load():
    if not os.path.exists(_LIB) or os.path.getmtime(_LIB) < os.path.getmtime(os.path.join(HERE, "exl2_cpu.c")):
        subprocess.run(["make", "-s", "-C", HERE], check=True)
    lib = ctypes.CDLL(_LIB)
    lib.exl2_cpu_threads.restype = ctypes.c_int
    return lib


def _p(a):
    return a.ctypes.data_as(ctypes.c_void_p) if a is not None else None


def exl2_args(w: dict, prescale: float = 1.0):
    """Checkpoint dict -> contiguous arrays in the C function's argument order (done once, outside any timing)."""
    import exl2_oracle as o
    qw = np.ascontiguousarray(w["q_weight"]).view(np.uint32)
    qs = np.ascontiguousarray(w["q_scale"]).view(np.uint32)
    smax = np.ascontiguousarray(o.prescale_max(w["q_scale_max"], prescale)).view(np.uint16)
    qg = np.ascontiguousarray(w["q_groups"]).astype(np.int16)
    K = int(w["q_invperm"].shape[0])
    perm = np.argsort(np.asarray(w["q_invperm"]).astype(np.int64), kind="stable").astype(np.uint16)
    return dict(qw=qw, qs=qs, smax=smax, qg=qg, perm=perm, K=K, N=qw.shape[1], G=qs.shape[0], R=qw.shape[0])


def exl2_gemv_prepared(lib, p: dict, a32: np.ndarray, y: np.ndarray):
    lib.exl2_cpu_gemv(_p(p["qw"]), _p(p["qs"]), _p(p["smax"]), _p(p["qg"]), _p(p["perm"]), p["K"], p["N"], p["G"], p["R"], _p(a32), _p(y))


def exl2_gemv(lib, w: dict, a: np.ndarray, prescale: float = 1.0) -> np.ndarray:
    p = exl2_args(w, prescale)
    a32 = np.ascontiguousarray(a, dtype=np.float32).reshape(p["K"])
    y = np.empty((p["N"],), dtype=np.float32)
    exl2_gemv_prepared(lib, p, a32, y)
    return y


def gptq_gemv(lib, w: dict, a: np.ndarray) -> np.ndarray:
    qw = np.ascontiguousarray(w["qweight"]).view(np.uint32)
    qz = np.ascontiguousarray(w["qzeros"]).view(np.uint32)
    sc = np.ascontiguousarray(w["scales"]).view(np.uint16)
    K, N = qw.shape[0] * 8, qw.shape[1]
    a32 = np.ascontiguousarray(a, dtype=np.float32).reshape(K)
    y = np.empty((N,), dtype=np.float32)
    lib.gptq_cpu_gemv(_p(qw), _p(qz), _p(sc), K, N, qz.shape[0], _p(a32), _p(y))
    return y
