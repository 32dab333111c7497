"""Decode-step plumbing around the quantized GEMVs: thin torch-tensor wrappers over the
`tg_decode_*` entry points of the C ABI (csrc/decode_ops.cu).  Single token, bf16 / fp16, CUDA only; launches go
to the current stream and are CUDA-graph capturable.  There is no CPU fallback.

The reference has no counterpart: its benchmark.py:145-146 times the stock HF model around the tinygemm ops.
"""
import ctypes

import torch

from . import _native

_DT = {torch.bfloat16: 0, torch.float16: 1}


def _p(t):
    return ctypes.c_void_p(t.data_ptr()) if t is not None else None


def _stream():
    return ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)


def _check(rc):
    if rc != 0:
        raise RuntimeError(_native.last_error())


def _req(t, name, like=None):
    if not (t.is_cuda and t.is_contiguous() and t.dtype in _DT):
        raise RuntimeError(f"{name} must be a contiguous CUDA bf16/fp16 tensor")
    if like is not None and (t.device != like.device or t.dtype != like.dtype):
        raise RuntimeError(f"{name} must live on {like.device} with dtype {like.dtype}")


def _one_token(t, name):
    if t.dim() > 1 and t.numel() != t.shape[-1]:
        raise RuntimeError(f"{name}: the decode ops take ONE token ({tuple(t.shape)} has more than one row)")


def _out(out, shape, like, name="out"):
    if out is None:
        return torch.empty(shape, device=like.device, dtype=like.dtype)
    _req(out, name, like)
    if tuple(out.shape) != tuple(shape):
        raise RuntimeError(f"{name} must have shape {tuple(shape)}")
    return out


def add_rmsnorm(h, delta, weight, eps, out=None):
    """h += delta (in place, rounded; delta may be None); returns rmsnorm(h, eps) * weight."""
    _req(h, "h"), _req(weight, "weight", h), _one_token(h, "h")
    if delta is not None:
        _req(delta, "delta", h)
        if delta.numel() != h.numel():
            raise RuntimeError("delta must match h")
    if weight.numel() != h.numel():
        raise RuntimeError("weight must match h")
    out = _out(out, h.shape, h)
    with torch.cuda.device(h.device):
        _check(_native.capi().tg_decode_add_rmsnorm(_p(h), _p(delta), _p(weight), _p(out), h.numel(), float(eps),
                                                    _DT[h.dtype], _stream()))
    return out


def silu_mul(gate_up, out=None):
    """gate_up = [gate | up] (2n values, e.g. the output of a fused gate/up GEMV) -> silu(gate) * up, n values."""
    _req(gate_up, "gate_up"), _one_token(gate_up, "gate_up")
    n = gate_up.numel() // 2
    out = _out(out, gate_up.shape[:-1] + (n,), gate_up)
    with torch.cuda.device(gate_up.device):
        _check(_native.capi().tg_decode_silu_mul(_p(gate_up), _p(out), n, _DT[gate_up.dtype], _stream()))
    return out


def rope_attention(qkv, cos, sin, k_cache, v_cache, pos, n_heads, n_kv_heads, head_dim=128, scale=None, out=None):
    """qkv = [q | k | v] of ONE token: rotary embedding of q and k, append k, v at `pos` to the caches
    [n_kv_heads][cache_len][head_dim], attention over positions 0..pos -> [n_heads * head_dim]."""
    for t, nm in ((qkv, "qkv"), (cos, "cos"), (sin, "sin"), (k_cache, "k_cache"), (v_cache, "v_cache")):
        _req(t, nm, qkv)
    _one_token(qkv, "qkv")
    if qkv.numel() != (n_heads + 2 * n_kv_heads) * head_dim:
        raise RuntimeError("qkv has the wrong size")
    if cos.numel() != head_dim or sin.numel() != head_dim:
        raise RuntimeError("cos / sin must have head_dim elements")
    cache_len = k_cache.numel() // (n_kv_heads * head_dim)
    if v_cache.numel() != k_cache.numel() or k_cache.numel() != n_kv_heads * cache_len * head_dim:
        raise RuntimeError("k_cache / v_cache must be [n_kv_heads][cache_len][head_dim]")
    scale = head_dim ** -0.5 if scale is None else scale
    out = _out(out, qkv.shape[:-1] + (n_heads * head_dim,), qkv)
    with torch.cuda.device(qkv.device):
        _check(_native.capi().tg_decode_rope_attention(_p(qkv), _p(cos), _p(sin), _p(k_cache), _p(v_cache), _p(out),
                                                       n_heads, n_kv_heads, head_dim, int(pos), cache_len, float(scale),
                                                       _DT[qkv.dtype], _stream()))
    return out


def linear_silu_pairs(lin, x):
    from .modules import RowShardedLinear

    if isinstance(lin, RowShardedLinear):  # row-sharded: activation AND exchange inside the GEMV kernel
        return lin.forward_silu_pairs(x)
    return _linear_silu_pairs_local(lin, x)


def _linear_silu_pairs_local(lin, x):
    """`lin`: a packed Any4Linear (weight-on-the-right kernel, per-row LUT, no bias) whose rows interleave a gate and
    an up projection (row 2j = gate_j, row 2j+1 = up_j, see modules.fuse_rows(..., interleave=True)); returns
    silu(gate(x)) * up(x) [m][out_features / 2] from ONE launch (activation fused into the GEMV epilogue)."""
    from .modules import Any4Linear

    if not (isinstance(lin, Any4Linear) and lin.weight_reshaped and lin.per_row and lin.bias is None
            and lin.kernel == "linear_y_f16RM_x_f16RM_W_any4TC"):
        raise RuntimeError("linear_silu_pairs needs a packed per-row-LUT Any4Linear with the weight on the right, no bias")
    _req(x, "x")
    if x.device != lin.weight.device:
        raise RuntimeError("linear_silu_pairs: x and the layer live on different devices")
    x2 = x.view(-1, x.shape[-1])
    w = lin.weight
    w_rows, ik = w.shape[0] * 8, w.shape[3] * 2
    y = torch.empty(x2.shape[0], w_rows // 2, device=x.device, dtype=x.dtype)
    with torch.cuda.device(x.device):
        _check(_native.capi().tg_gemm_w4_rm_silu_pairs(_p(y), _p(x2), _p(w), _p(lin.scales_and_zeros), _p(lin.lut), None,
                                                       x2.shape[0], w_rows, x2.shape[1], lin.group_size, ik, 2,
                                                       _DT[x.dtype], _stream()))
    return y.view(*x.shape[:-1], w_rows // 2)
