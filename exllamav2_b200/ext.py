"""`exllamav2_ext`-compatible operator surface over libexl2b200.so (ctypes, C ABI in include/exl2_b200.h).

Every public function below has the NAME, ARGUMENT ORDER and meaning of the reference pybind binding it replaces
(exllamav2/exllamav2_ext/ext_bindings.cpp:27-138), so exllamav2/{linear,attn,mlp,cache,rmsnorm}.py call it
unchanged when this module is importable as `exllamav2_ext` (`install_as_exllamav2_ext()`).
torch is used only for device memory / the current stream -- all arithmetic is in the CUDA library.

The product path fails loudly: importing this module without the built library, or calling an op on CPU tensors,
raises.  There is no CPU or PyTorch fallback.
"""
from __future__ import annotations

import ctypes
import os
import sys
from ctypes import POINTER, Structure, c_char_p, c_float, c_int, c_int16, c_int32, c_uint64, c_void_p

import torch

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.path.join(_HERE, "libexl2b200.so")

if not os.path.exists(_LIB_PATH):
    raise ImportError(
        f"{_LIB_PATH} is missing: build it with `python -m exllamav2_b200.build` (nvcc, sm_100a). "
        "exllamav2_b200 has no CPU fallback.")

lib = ctypes.CDLL(_LIB_PATH)


class _QMatrixDesc(Structure):
    _fields_ = [
        ("device", c_int), ("height", c_int), ("width", c_int), ("groups", c_int),
        ("q_weight", c_void_p), ("q_perm", c_void_p), ("q_invperm", c_void_p),
        ("q_scale", c_void_p), ("q_scale_max", c_void_p), ("q_groups", c_void_p), ("q_weight_rows", c_int),
        ("gptq_qzeros", c_void_p), ("gptq_scales", c_void_p), ("gptq_g_idx", c_void_p), ("bias", c_void_p),
    ]


class _QAttnDesc(Structure):
    _fields_ = [
        ("layernorm", c_void_p), ("norm_epsilon", c_float),
        ("q_proj", c_void_p), ("k_proj", c_void_p), ("v_proj", c_void_p), ("o_proj", c_void_p),
        ("hidden_size", c_int), ("num_heads", c_int), ("num_kv_heads", c_int), ("head_dim", c_int),
        ("has_residual", c_int), ("rope_style", c_int), ("sincos_size", c_int),
    ]


class _QMlpDesc(Structure):
    _fields_ = [
        ("layernorm", c_void_p), ("norm_epsilon", c_float),
        ("gate", c_void_p), ("up", c_void_p), ("down", c_void_p),
        ("hidden_size", c_int), ("intermediate_size", c_int), ("act_gelu", c_int), ("has_residual", c_int),
    ]


def _sig(name, restype, *argtypes):
    f = getattr(lib, name)
    f.restype = restype
    f.argtypes = list(argtypes)
    return f


_sig("exl2b_last_error", c_char_p)
_sig("exl2b_version", c_int)
_sig("exl2b_launch_count", c_uint64)
_sig("exl2b_qmatrix_create", c_int, POINTER(_QMatrixDesc), c_void_p, POINTER(c_void_p))
_sig("exl2b_qmatrix_destroy", c_int, c_void_p)
_sig("exl2b_qmatrix_info", c_int, c_void_p, POINTER(c_int), POINTER(c_int), POINTER(c_int), POINTER(c_int), POINTER(c_uint64))
_sig("exl2b_reconstruct", c_int, c_void_p, c_void_p, c_void_p)
_sig("exl2b_gemm_half_q_half", c_int, c_void_p, c_void_p, c_int, c_void_p, c_int, c_int, c_int, c_int, c_void_p)
_sig("exl2b_gemm_half_q_half_norm", c_int, c_void_p, c_void_p, c_void_p, c_float, c_void_p, c_int, c_void_p)
_sig("exl2b_gemm_half_q_half_host", c_int, c_void_p, c_void_p, c_void_p, c_int, c_void_p)
_sig("exl2b_make_group_map", c_int, POINTER(c_int16), c_int, c_int, POINTER(c_int16), c_int, POINTER(c_int))
_sig("exl2b_rms_norm", c_int, c_void_p, c_void_p, c_void_p, c_float, c_int, c_int, c_void_p)
_sig("exl2b_rope", c_int, c_void_p, c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_int, c_void_p, c_int, c_int, c_void_p)
_sig("exl2b_act_mul", c_int, c_void_p, c_void_p, c_int, c_int, c_int, c_void_p)
_KV_ARGS = (c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_int, c_int,
            c_void_p, c_void_p, c_int, c_int, c_void_p)
_sig("exl2b_fp16_to_q_kv", c_int, *_KV_ARGS)
_sig("exl2b_q_to_fp16_kv", c_int, *_KV_ARGS)
_sig("exl2b_qattn_create", c_int, POINTER(_QAttnDesc), POINTER(c_void_p))
_sig("exl2b_qattn_destroy", c_int, c_void_p)
_sig("exl2b_qattn_forward_1", c_int, c_void_p, c_void_p, c_int, c_int, c_int, c_void_p, c_void_p, c_void_p, c_void_p,
     c_void_p, c_void_p, c_void_p)
_sig("exl2b_qattn_forward_2", c_int, c_void_p, c_void_p, c_void_p, c_int, c_int, c_void_p)
_sig("exl2b_qmlp_create", c_int, POINTER(_QMlpDesc), POINTER(c_void_p))
_sig("exl2b_qmlp_destroy", c_int, c_void_p)
_sig("exl2b_qmlp_forward", c_int, c_void_p, c_void_p, c_int, c_void_p, c_void_p, c_void_p)
_sig("exl2b_qmlp_forward_gateup", c_int, c_void_p, c_void_p, c_int, c_void_p, c_void_p)
_sig("exl2b_paged_attn_decode_q4", c_int, *([c_void_p] * 10 + [c_int] * 7 + [c_float, c_void_p, c_void_p]))
_sig("exl2b_paged_attn_decode_q4_ex", c_int, *([c_void_p] * 10 + [c_int] * 7 + [c_float, c_void_p, c_void_p, c_void_p, c_int, c_int, c_void_p]))
_sig("exl2b_paged_attn_status", c_int, c_int, POINTER(c_int))
_sig("exl2b_paged_attn_clear_status", c_int, c_int)


class _Chain(Structure):
    _fields_ = [("consumers", c_void_p * 3), ("num_consumers", c_int), ("norm_weight", c_void_p)]


_sig("exl2b_qattn_forward_1_ex", c_int, c_void_p, c_void_p, c_int, c_int, c_int, c_void_p, c_void_p, c_void_p, c_void_p,
     c_void_p, c_void_p, c_int, c_void_p)
_sig("exl2b_qattn_forward_2_ex", c_int, c_void_p, c_void_p, c_void_p, c_int, c_int, c_int, POINTER(_Chain), c_void_p)
_sig("exl2b_qmlp_forward_ex", c_int, c_void_p, c_void_p, c_int, c_void_p, c_void_p, c_int, POINTER(_Chain), c_void_p)
_sig("exl2b_gemm_half_q_half_prepared", c_int, c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_float, c_void_p)

# Dummy tensor standing for None/NULL (ext.py:296 of the reference)
none_tensor = torch.empty((1, 1), device="meta")


def _check(rc: int):
    if rc != 0:
        raise RuntimeError(lib.exl2b_last_error().decode())


def _p(t: torch.Tensor | None):
    """device pointer or NULL for the meta-device `none_tensor`."""
    if t is None or t.is_meta:
        return None
    return t.data_ptr()


def _cuda(t: torch.Tensor, what: str):
    if not t.is_cuda:
        raise RuntimeError(f"{what} must be a CUDA tensor: exllamav2_b200 has no CPU path")
    return t


def _dtype(t: torch.Tensor, dt, what: str):
    if t.dtype != dt:
        raise RuntimeError(f"{what} is incorrect datatype, must be {dt}")     # TORCH_CHECK_DTYPE, cpp/util.h:34


def _stream(t: torch.Tensor):
    return torch.cuda.current_stream(t.device).cuda_stream


def launch_count() -> int:
    return int(lib.exl2b_launch_count())


# ---------------------------------------------------------------------------------------------------------------
# QMatrix
# ---------------------------------------------------------------------------------------------------------------

def make_q_matrix(q_weight, q_perm, q_invperm, q_scale, q_scale_max, q_groups, q_group_map, gptq_qzeros, gptq_scales,
                  gptq_g_idx, bias, temp_dq, max_dq_rows: int) -> int:
    """ext_qmatrix.cpp:21-111.  Same 13 arguments; temp_dq / max_dq_rows / q_group_map are accepted and unused
    (no reconstruct+cuBLAS detour, the group table is rebuilt from q_groups)."""
    _cuda(q_weight, "q_weight")
    _dtype(q_weight, torch.int32, "q_weight")
    d = _QMatrixDesc()
    d.device = q_weight.device.index or 0
    d.width = q_weight.shape[1]
    d.q_weight = _p(q_weight)
    d.q_weight_rows = q_weight.shape[0]
    d.q_perm = _p(q_perm)
    d.q_invperm = _p(q_invperm)
    d.bias = _p(bias)
    if not q_scale.is_meta:
        _dtype(q_scale, torch.int32, "q_scale")
        _dtype(q_scale_max, torch.float16, "q_scale_max")
        _dtype(q_groups, torch.int16, "q_groups")
        if q_weight.shape[1] != q_scale.shape[1] * 8:
            raise RuntimeError("q_weight and q_scale have incompatible shapes")      # TORCH_CHECK_SHAPES(...,8)
        d.groups = q_scale.shape[0]
        if not q_perm.is_meta:
            _dtype(q_perm, torch.int16, "q_perm")
            d.height = q_perm.shape[0]
        elif not q_group_map.is_meta:
            d.height = q_group_map.shape[0] // 2
        else:
            raise RuntimeError("EXL2 matrix needs q_perm or q_group_map to define its height")
        d.q_scale = _p(q_scale)
        d.q_scale_max = _p(q_scale_max)
        d.q_groups = _p(q_groups)
        keep = None
    else:
        _dtype(gptq_qzeros, torch.int32, "gptq_qzeros")
        _dtype(gptq_scales, torch.float16, "gptq_scales")
        if q_weight.shape[1] != gptq_qzeros.shape[1] * 8 or q_weight.shape[1] != gptq_scales.shape[1]:
            raise RuntimeError("qweight, qzeros and scales have incompatible shapes")
        d.groups = gptq_qzeros.shape[0]
        d.height = q_weight.shape[0] * 8
        d.gptq_qzeros = _p(gptq_qzeros)
        d.gptq_scales = _p(gptq_scales)
        keep = None
        if not gptq_g_idx.is_meta:
            keep = gptq_g_idx.to(device="cpu", dtype=torch.int32).contiguous()
            d.gptq_g_idx = keep.data_ptr()
    out = c_void_p()
    with torch.cuda.device(q_weight.device):
        _check(lib.exl2b_qmatrix_create(ctypes.byref(d), _stream(q_weight), ctypes.byref(out)))
    del keep
    return out.value


def free_q_matrix(handle: int):
    _check(lib.exl2b_qmatrix_destroy(handle))


def q_matrix_info(handle: int) -> dict:
    h, w, g, gq, pb = c_int(), c_int(), c_int(), c_int(), c_uint64()
    _check(lib.exl2b_qmatrix_info(handle, ctypes.byref(h), ctypes.byref(w), ctypes.byref(g), ctypes.byref(gq), ctypes.byref(pb)))
    return {"height": h.value, "width": w.value, "groups": g.value, "is_gptq": bool(gq.value), "packed_bytes": pb.value}


def reconstruct(q_handle: int, output: torch.Tensor):
    """ext_qmatrix.cpp:196-210"""
    _cuda(output, "output")
    _dtype(output, torch.float16, "output")
    info = q_matrix_info(q_handle)
    if info["height"] != output.shape[0] or info["width"] != output.shape[1]:
        raise RuntimeError("Output tensor doesn't match shape of QMatrix")
    _check(lib.exl2b_reconstruct(q_handle, output.data_ptr(), _stream(output)))


def gemm_half_q_half(a: torch.Tensor, b: int, c: torch.Tensor, force_cuda: bool = False):
    """ext_qmatrix.cpp:213-247: c = a @ W (+ bias), a fp16[M,K], c fp16[M,N]."""
    _cuda(a, "a")
    _dtype(a, torch.float16, "a")
    _dtype(c, torch.float16, "c")
    if a.shape[0] != c.shape[0]:
        raise RuntimeError("a and c have incompatible shapes")
    info = q_matrix_info(b)
    if info["height"] != a.shape[1]:
        raise RuntimeError("a and b have incompatible shapes")
    if info["width"] != c.shape[1]:
        raise RuntimeError("b and c have incompatible shapes")
    if a.stride(1) != 1 or c.stride(1) != 1:
        raise RuntimeError("a and c must be row-major")
    _check(lib.exl2b_gemm_half_q_half(b, a.data_ptr(), a.stride(0), c.data_ptr(), c.stride(0), a.shape[0], 1,
                                      int(force_cuda), _stream(a)))


def gemv_norm(x: torch.Tensor, b: int, w: torch.Tensor, epsilon: float, c: torch.Tensor, clear: bool = True, prepared: bool = False):
    """rms_norm(x, w) followed by gemm_half_q_half, fused into one launch for a single row (decode: final norm + lm_head);
    more rows run the two reference ops (rmsnorm.py:141, linear.py:366)."""
    _cuda(x, "x")
    _dtype(x, torch.float16, "x")
    _dtype(c, torch.float16, "c")
    rows = x.numel() // x.shape[-1]
    if rows == 1:
        # prepared: the row was left in the matrix's stored-row order by a chained producer launch (exl2b_chain_t)
        _check(lib.exl2b_gemm_half_q_half_norm(b, None if prepared else x.data_ptr(), w.data_ptr(), float(epsilon), c.data_ptr(),
                                               int(clear), _stream(x)))
        return
    y = torch.empty_like(x)
    rms_norm(x, w, y, epsilon)
    if clear:
        gemm_half_q_half(y.view(rows, -1), b, c.view(rows, -1), False)
    else:
        gemm_half_q_half_accum(y.view(rows, -1), b, c.view(rows, -1))


def gemm_half_q_half_accum(a: torch.Tensor, b: int, c: torch.Tensor):
    """c += a @ W -- the clear=false form the reference uses internally for residual adds (cuda/q_attn.cu:333)."""
    _check(lib.exl2b_gemm_half_q_half(b, a.data_ptr(), a.stride(0), c.data_ptr(), c.stride(0), a.shape[0], 0, 0, _stream(a)))


def gemm_half_q_half_host(a_host: torch.Tensor, b: int, c_host: torch.Tensor, device: torch.device):
    """Host-buffer entry point (bench.py e2e): copies a to the device, multiplies, copies c back, waits."""
    _check(lib.exl2b_gemm_half_q_half_host(b, a_host.data_ptr(), c_host.data_ptr(), a_host.shape[0],
                                           torch.cuda.current_stream(device).cuda_stream))


def make_group_map(q_groups: torch.Tensor, num_qrows: int) -> torch.Tensor:
    """ext_qmatrix.cpp:341-361 (CPU tensor in, CPU int16 tensor out)."""
    _dtype(q_groups, torch.int16, "q_groups")
    g = q_groups.cpu().contiguous()
    ng = g.shape[0] // 2
    bits = g[0::2].to(torch.int64)
    cap = int((num_qrows * 32 // max(1, int(bits.min()))) * 2 + 64)
    out = torch.empty((cap,), dtype=torch.int16)
    k = c_int()
    _check(lib.exl2b_make_group_map(ctypes.cast(g.data_ptr(), POINTER(c_int16)), ng, num_qrows,
                                    ctypes.cast(out.data_ptr(), POINTER(c_int16)), cap, ctypes.byref(k)))
    return out[: 2 * k.value].clone()


# ---------------------------------------------------------------------------------------------------------------
# norm / rope / activation
# ---------------------------------------------------------------------------------------------------------------

def rms_norm(x: torch.Tensor, w: torch.Tensor, y: torch.Tensor, epsilon: float):
    """ext_norm.cpp:23-62 (fp16 in / fp16 out form)."""
    _cuda(x, "x")
    _dtype(x, torch.float16, "x")
    _dtype(w, torch.float16, "w")
    _dtype(y, torch.float16, "y")
    dim = x.shape[-1]
    if w.shape[0] != dim:
        raise RuntimeError("x and w have incompatible shapes")
    rows = x.numel() // dim
    _check(lib.exl2b_rms_norm(x.data_ptr(), w.data_ptr(), y.data_ptr(), float(epsilon), rows, dim, _stream(x)))


def rms_norm_(x: torch.Tensor, w: torch.Tensor, epsilon: float):
    """ext_norm.cpp:64-103: in place."""
    rms_norm(x, w, x, epsilon)


def rope_(x: torch.Tensor, sin: torch.Tensor, cos: torch.Tensor, past_len: int, num_heads: int, head_dim: int,
          offsets: torch.Tensor, neox_style: bool):
    """ext_rope.cpp:20-62: x fp16[batch, seq, heads*head_dim] rotated in place."""
    _cuda(x, "x")
    _dtype(x, torch.float16, "x")
    _dtype(sin, torch.float16, "sin")
    _dtype(cos, torch.float16, "cos")
    if head_dim * num_heads != x.shape[-1]:
        raise RuntimeError("x has wrong last dimension for num_heads * head_dim")
    batch = x.shape[0]
    rows_per_batch = x.numel() // head_dim // batch
    sincos_size = sin.shape[-1]
    _check(lib.exl2b_rope(x.data_ptr(), sin.data_ptr(), cos.data_ptr(), batch, rows_per_batch, head_dim, num_heads,
                          int(past_len), _p(offsets), int(bool(neox_style)), sincos_size, _stream(x)))


def act_mul(x: torch.Tensor, y: torch.Tensor, act_gelu: bool = False):
    _check(lib.exl2b_act_mul(x.data_ptr(), y.data_ptr(), x.numel() // x.shape[-1], x.shape[-1], int(act_gelu), _stream(x)))


# ---------------------------------------------------------------------------------------------------------------
# Q4 K/V cache
# ---------------------------------------------------------------------------------------------------------------

def _kv_call(fn, a_k, b_k, s_k, a_v, b_v, s_v, fp16_k, batch_size, offset, width, page_size, cache_seqlens, block_table, wbits):
    dim = fp16_k.shape[2] * fp16_k.shape[3]
    seq_stride = fp16_k.shape[1] * dim
    pages = 0
    if page_size:
        batch_size = block_table.shape[0]
        pages = block_table.shape[1]
        if cache_seqlens.shape[0] != block_table.shape[0]:
            raise RuntimeError("cache_seqlens and block_table have incompatible shapes")
    _check(fn(a_k.data_ptr(), b_k.data_ptr(), s_k.data_ptr(), _p(a_v), _p(b_v), _p(s_v), batch_size, dim, seq_stride,
              offset, width, page_size, _p(cache_seqlens), _p(block_table), pages, wbits, _stream(a_k)))


def fp16_to_q_kv(k_in, k_out, k_scales, v_in, v_out, v_scales, batch_size: int, offset: int, width: int, page_size: int,
                 cache_seqlens, block_table, wbits: int):
    """ext_cache.cpp:80-174"""
    _cuda(k_in, "k_in")
    _dtype(k_in, torch.float16, "k_in")
    _dtype(k_out, torch.uint8, "k_out")
    _kv_call(lib.exl2b_fp16_to_q_kv, k_in, k_out, k_scales, v_in, v_out, v_scales, k_in, batch_size, offset, width,
             page_size, cache_seqlens, block_table, wbits)


def q_to_fp16_kv(k_in, k_out, k_scales, v_in, v_out, v_scales, batch_size: int, offset: int, width: int, page_size: int,
                 cache_seqlens, block_table, wbits: int):
    """ext_cache.cpp:176-274 (argument order of the reference: in, OUT, scales)."""
    _cuda(k_in, "k_in")
    _dtype(k_in, torch.uint8, "k_in")
    _dtype(k_out, torch.float16, "k_out")
    # C ABI order is (in, scales, out)
    dim = k_out.shape[2] * k_out.shape[3]
    seq_stride = k_out.shape[1] * dim
    pages = 0
    if page_size:
        batch_size = block_table.shape[0]
        pages = block_table.shape[1]
    _check(lib.exl2b_q_to_fp16_kv(k_in.data_ptr(), k_scales.data_ptr(), k_out.data_ptr(), _p(v_in), _p(v_scales), _p(v_out),
                                  batch_size, dim, seq_stride, offset, width, page_size, _p(cache_seqlens), _p(block_table),
                                  pages, wbits, _stream(k_in)))


# ---------------------------------------------------------------------------------------------------------------
# fused blocks
# ---------------------------------------------------------------------------------------------------------------

def make_q_attn(layernorm, layernorm_bias, layernorm_is_rms: bool, headnorm_is_rms: bool, norm_epsilon: float,
                q_q_proj: int, q_k_proj: int, q_v_proj: int, q_o_proj: int, temp_state, temp_dq, max_rows: int,
                hidden_size: int, num_heads: int, num_kv_heads: int, head_dim: int, max_seq_len: int, has_residual: bool,
                rope_style: int, sincos_size: int, q_norm, k_norm, post_layernorm, post_layernorm_bias,
                residual_fp32: bool, use_graphs: bool) -> int:
    """ext_qattn.cpp:24-104, same 26 arguments.  Llama-family subset: RMSNorm pre-norm, no QK-norm, no
    post-norm, fp16 residual stream; anything else raises (out of scope)."""
    if not layernorm.is_meta and not layernorm_is_rms:
        raise RuntimeError("exllamav2_b200: only RMSNorm pre-normalisation is implemented")
    for t, name in ((q_norm, "q_norm"), (k_norm, "k_norm"), (post_layernorm, "post_layernorm")):
        if t is not None and not t.is_meta:
            raise RuntimeError(f"exllamav2_b200: {name} is not implemented (out of scope)")
    if residual_fp32:
        raise RuntimeError("exllamav2_b200: fp32 residual stream is not implemented")
    d = _QAttnDesc()
    d.layernorm = _p(layernorm)
    d.norm_epsilon = float(norm_epsilon)
    d.q_proj, d.k_proj, d.v_proj, d.o_proj = q_q_proj, q_k_proj, q_v_proj, q_o_proj
    d.hidden_size, d.num_heads, d.num_kv_heads, d.head_dim = hidden_size, num_heads, num_kv_heads, head_dim
    d.has_residual = int(has_residual)
    d.rope_style = int(rope_style)
    d.sincos_size = int(sincos_size)
    out = c_void_p()
    _check(lib.exl2b_qattn_create(ctypes.byref(d), ctypes.byref(out)))
    return out.value


def free_q_attn(handle: int):
    _check(lib.exl2b_qattn_destroy(handle))


def q_attn_forward_1(q_attn: int, x, batch_size: int, q_len: int, past_len: int, past_lens, q_temp, k_temp, v_temp, sin, cos,
                     loras=(), loras_temp=none_tensor):
    """ext_qattn.cpp:115-159"""
    if loras:
        raise RuntimeError("exllamav2_b200: LoRA is out of scope")
    _cuda(x, "x")
    _dtype(x, torch.float16, "x")
    _check(lib.exl2b_qattn_forward_1(q_attn, x.data_ptr(), batch_size, q_len, int(past_len), _p(past_lens), q_temp.data_ptr(),
                                     k_temp.data_ptr(), v_temp.data_ptr(), _p(sin), _p(cos), _stream(x)))


def q_attn_forward_2(q_attn: int, x, attn_output, batch_size: int, q_len: int, loras=(), loras_temp=none_tensor):
    """ext_qattn.cpp:161-191"""
    if loras:
        raise RuntimeError("exllamav2_b200: LoRA is out of scope")
    _check(lib.exl2b_qattn_forward_2(q_attn, x.data_ptr(), attn_output.data_ptr(), batch_size, q_len, _stream(x)))


def make_q_mlp(layernorm, layernorm_bias, layernorm_is_rms: bool, norm_epsilon: float, q_gate: int, q_up: int, q_down: int,
               temp_state, temp_a, temp_b, temp_dq, max_rows: int, act_gelu: bool, has_residual: bool, post_layernorm,
               post_layernorm_bias, residual_fp32: bool, use_graphs: bool) -> int:
    """ext_qmlp.cpp:22-85, same arguments.  temp_a / temp_b are remembered per handle like the reference does."""
    if not layernorm.is_meta and not layernorm_is_rms:
        raise RuntimeError("exllamav2_b200: only RMSNorm pre-normalisation is implemented")
    if post_layernorm is not None and not post_layernorm.is_meta:
        raise RuntimeError("exllamav2_b200: post_layernorm is not implemented (out of scope)")
    if residual_fp32:
        raise RuntimeError("exllamav2_b200: fp32 residual stream is not implemented")
    if not q_gate:
        raise RuntimeError("exllamav2_b200: un-gated MLP is not implemented (out of scope)")
    gi, ui = q_matrix_info(q_gate), q_matrix_info(q_up)
    d = _QMlpDesc()
    d.layernorm = _p(layernorm)
    d.norm_epsilon = float(norm_epsilon)
    d.gate, d.up, d.down = q_gate, q_up, q_down
    d.hidden_size = gi["height"]
    d.intermediate_size = ui["width"]
    d.act_gelu = int(act_gelu)
    d.has_residual = int(has_residual)
    out = c_void_p()
    _check(lib.exl2b_qmlp_create(ctypes.byref(d), ctypes.byref(out)))
    _mlp_temps[out.value] = (temp_a, temp_b)
    return out.value


_mlp_temps: dict[int, tuple] = {}


def q_mlp_forward_rows(q_mlp: int, x, temp_a, temp_b):
    """q_mlp_forward_ with caller-provided scratch rows (the reference sizes temp_a / temp_b for max_input_len at make_q_mlp
    time, mlp.py:176-203; a prompt chunk larger than that brings its own)."""
    _cuda(x, "x")
    _dtype(x, torch.float16, "x")
    rows = x.numel() // x.shape[-1]
    _check(lib.exl2b_qmlp_forward(q_mlp, x.data_ptr(), rows, temp_a.data_ptr(), temp_b.data_ptr(), _stream(x)))


def free_q_mlp(handle: int):
    _mlp_temps.pop(handle, None)
    _check(lib.exl2b_qmlp_destroy(handle))


def q_mlp_forward_(q_mlp: int, x, loras=(), loras_temp=none_tensor):
    """ext_qmlp.cpp:87-118: x is updated in place."""
    if loras:
        raise RuntimeError("exllamav2_b200: LoRA is out of scope")
    _cuda(x, "x")
    _dtype(x, torch.float16, "x")
    temp_a, temp_b = _mlp_temps[q_mlp]
    rows = x.numel() // x.shape[-1]
    if rows > temp_a.shape[0]:          # the reference sizes temp_a / temp_b for max_rows at make_q_mlp time and would write past them
        raise RuntimeError(f"q_mlp_forward_: {rows} rows exceed the {temp_a.shape[0]} rows of temp_a given to make_q_mlp "
                           "(use q_mlp_forward_rows with scratch of your own for larger chunks)")
    _check(lib.exl2b_qmlp_forward(q_mlp, x.data_ptr(), rows, temp_a.data_ptr(), _p(temp_b), _stream(x)))


# names the reference's hot-path call sites use that this module provides

def q_mlp_forward_gateup(q_mlp: int, x, temp_a):
    """temp_a[rows, intermediate] = act(norm(x) @ gate) * (norm(x) @ up) -- the first half of q_mlp_forward_, used by a
    tensor-parallel rank on its intermediate slice (ext_qmlp.cpp:326-473)."""
    rows = x.numel() // x.shape[-1]
    _check(lib.exl2b_qmlp_forward_gateup(q_mlp, x.data_ptr(), rows, temp_a.data_ptr(), _stream(x)))


def make_chain(consumers, norm_weight=None) -> "_Chain":
    """Describe who reads a launch's output: `consumers` = q_handles of the matrices fed by it, `norm_weight` = the
    RMSNorm weight they apply (include/exl2_b200.h exl2b_chain_t).  Keep the returned object (and norm_weight) alive."""
    c = _Chain()
    for i, hnd in enumerate(consumers):
        c.consumers[i] = hnd
    c.num_consumers = len(consumers)
    c.norm_weight = _p(norm_weight)
    c._keep = norm_weight
    return c


def q_attn_forward_1_ex(q_attn: int, x, batch_size: int, q_len: int, past_len: int, past_lens, q_temp, k_temp, v_temp, sin, cos,
                        input_prepared: bool):
    _check(lib.exl2b_qattn_forward_1_ex(q_attn, _p(x), batch_size, q_len, int(past_len), _p(past_lens), q_temp.data_ptr(),
                                        k_temp.data_ptr(), v_temp.data_ptr(), _p(sin), _p(cos), int(input_prepared), _stream(q_temp)))


def q_attn_forward_2_ex(q_attn: int, x, attn_output, batch_size: int, q_len: int, input_prepared: bool, chain=None):
    _check(lib.exl2b_qattn_forward_2_ex(q_attn, x.data_ptr(), _p(attn_output), batch_size, q_len, int(input_prepared),
                                        ctypes.byref(chain) if chain is not None else None, _stream(x)))


def q_mlp_forward_ex(q_mlp: int, x, input_prepared: bool, chain=None):
    temp_a, temp_b = _mlp_temps[q_mlp]
    rows = x.numel() // x.shape[-1]
    _check(lib.exl2b_qmlp_forward_ex(q_mlp, x.data_ptr(), rows, temp_a.data_ptr(), _p(temp_b), int(input_prepared),
                                     ctypes.byref(chain) if chain is not None else None, _stream(x)))


def gemm_half_q_half_prepared(b: int, c, has_norm: bool, norm_eps: float, clear: bool = True):
    _check(lib.exl2b_gemm_half_q_half_prepared(b, c.data_ptr(), c.stride(0), c.shape[0], int(clear), int(has_norm),
                                               float(norm_eps), _stream(c)))


def paged_attn_decode_q4(q, k_new, v_new, k_cache, k_scales, v_cache, v_scales, cache_seqlens, block_table, out,
                         softmax_scale: float, out_consumer: int = 0, rope=None):
    """Decode attention over the paged Q4 cache with quantise-and-append of the new rows (include/exl2_b200.h
    exl2b_paged_attn_decode_q4).  q [B, q_len, H, hd]; k_new / v_new [B, q_len, KVH, hd]; caches uint8
    [pages, page, KVH, hd/2] + fp16 scales [pages, page, KVH, hd/32]; out like q.  No reference counterpart as one op:
    it replaces q_to_fp16_kv + flash_attn_with_kvcache + fp16_to_q_kv (attn.py:560-613)."""
    B, q_len, H, hd = q.shape
    KVH = k_new.shape[2]
    for t in (q, k_new, v_new, out):
        _dtype(_cuda(t, "attention operand"), torch.half, "attention operand")
    # rope = (sin, cos, rope_style): q / k_new are the un-rotated projection outputs, rotated as the kernel reads them
    sin, cos, style = rope if rope is not None else (None, None, 0)
    _check(lib.exl2b_paged_attn_decode_q4_ex(
        _p(q), _p(k_new), _p(v_new), _p(k_cache), _p(k_scales), _p(v_cache), _p(v_scales),
        _p(cache_seqlens), _p(block_table), _p(out), B, q_len, H, KVH, hd, k_cache.shape[1], block_table.shape[1],
        float(softmax_scale), out_consumer or None, _p(sin), _p(cos), int(style), sin.shape[-1] if sin is not None else 0,
        _stream(q)))


def paged_attn_clear_status(device) -> None:
    _check(lib.exl2b_paged_attn_clear_status(torch.device(device).index or 0))


def paged_attn_status(device) -> int:
    """Sticky error bits of the fused attention kernels (bit 0: a sequence ran past its page table).  Synchronises."""
    st = c_int(0)
    _check(lib.exl2b_paged_attn_status(torch.device(device).index or 0, ctypes.byref(st)))
    return st.value


HOT_PATH_EXPORTS = [
    "make_q_matrix", "free_q_matrix", "reconstruct", "gemm_half_q_half", "make_group_map",
    "rms_norm", "rms_norm_", "rope_", "fp16_to_q_kv", "q_to_fp16_kv",
    "make_q_attn", "free_q_attn", "q_attn_forward_1", "q_attn_forward_2",
    "make_q_mlp", "free_q_mlp", "q_mlp_forward_",
]


# Names of the reference extension that are NOT on the hot path (sampling, safetensors loader, MoE, LoRA, head / layer norm,
# FP8 cache, converter kernels, the single-process TP glue -- SURVEY.md 2.2) are forwarded to the stock extension when the
# deployment registers one; otherwise the AttributeError says exactly which name is missing and why.
_stock_ext = None


def set_stock_extension(module) -> None:
    """Register the reference's own build of `exllamav2_ext` (or any module exporting its names) as the provider of
    everything outside the hot path."""
    global _stock_ext
    _stock_ext = module


def __getattr__(name: str):
    if name.startswith("__"):
        raise AttributeError(name)
    stock = _stock_ext
    if stock is None:
        try:
            import importlib
            stock = importlib.import_module("exllamav2_ext_stock")      # a deployment may put the stock build on sys.path under this name
            set_stock_extension(stock)
        except ImportError:
            stock = None
    if stock is not None and hasattr(stock, name):
        return getattr(stock, name)
    raise AttributeError(f"exllamav2_b200.ext: '{name}' is outside the quantized-linear hot path this module replaces and no stock "
                         f"exllamav2_ext is registered (exllamav2_b200.ext.set_stock_extension / module 'exllamav2_ext_stock')")


def install_as_exllamav2_ext():
    """Register this module under the name the reference imports (exllamav2/ext.py:106-109 does
    `import exllamav2_ext` first and only JIT-builds its own extension when that fails)."""
    sys.modules["exllamav2_ext"] = sys.modules[__name__]
