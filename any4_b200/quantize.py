"""any4 quantizer front-end on the GPU: float weight -> (codes, any4 table, scales_and_zeros) or
straight to a packed `Any4Linear`, one kernel launch per weight matrix (include/tinygemm_b200.h: tg_quantize_any4_rows).

Mirrors the call the reference makes for the any4 Linear (quantize.py:523-610 `anyq_quantize_tensor` with n_bit = 4,
per_row = True, zero_point = True, init = "int"; quantize.py:880-900 builds the module from its outputs): same names,
same return values, same layouts; the clustering is the reference's own Lloyd iteration (kmeans.py:230-287) run on the
device instead of sklearn / joblib on the host.  There is no CPU fallback.
"""
import ctypes

import torch

from . import _native

_DT = {torch.bfloat16: 0, torch.float16: 1}


def _p(t):
    return None if t is None else ctypes.c_void_p(t.data_ptr())


def anyq_quantize_tensor(W, n_bit=4, q_group_size=128, per_row=True, zero_point=True, init="int", sample_weight=None,
                         max_iter=300, tol=1e-4, pack_inner_k=None, return_lut=False):
    """W [n][k] bf16 / fp16 on a CUDA device ->  assign [n][k] int32, any4 [n][16] W.dtype (code space [0, 15], NOT
    centred), scales_and_zeros [k/g][n][2] W.dtype  (quantize.py:523-610).  `sample_weight`: optional [k] weights of
    the k-means (e.g. mean squared activations).  `pack_inner_k` = 2 / 4 / 8: also return the weight already in the B
    int4 tensor-core layout (what `reshape_weight` would produce from `assign`); `return_lut`: also `any4 - 8`."""
    if n_bit != 4 or not per_row or not zero_point or init != "int":
        raise NotImplementedError("the GPU front-end covers the any4 Linear's configuration: 4 bit, per-row table, "
                                  "asymmetric groups with zero point, init='int'")
    if not (W.is_cuda and W.dim() == 2 and W.dtype in _DT):
        raise RuntimeError("W must be a 2-D CUDA bf16/fp16 tensor")
    W = W.contiguous()
    n, k = W.shape
    dev, dt = W.device, W.dtype
    sw = None
    if sample_weight is not None:
        sw = sample_weight.to(device=dev, dtype=torch.float32).abs().contiguous()  # build_sample_weight(abs=True), kmeans.py:136
        if sw.numel() != k:
            raise RuntimeError("sample_weight must have one entry per input feature")
    assign = torch.empty((n, k), device=dev, dtype=torch.int32)
    any4 = torch.empty((n, 16), device=dev, dtype=dt)
    lut = torch.empty((n, 16), device=dev, dtype=dt)
    sz = torch.empty((k // q_group_size, n, 2), device=dev, dtype=dt)
    packed = None
    if pack_inner_k is not None:
        packed = torch.empty((n // 8, -(-(k // 16) // pack_inner_k), 32, pack_inner_k // 2), device=dev, dtype=torch.int32)
    with torch.cuda.device(dev):
        rc = _native.capi().tg_quantize_any4_rows(
            _p(W), _p(sw), n, k, q_group_size, pack_inner_k or 0, int(max_iter), float(tol), _p(assign), _p(packed), _p(sz),
            _p(any4), _p(lut), None, _DT[dt], ctypes.c_void_p(torch.cuda.current_stream().cuda_stream))
    if rc != 0:
        raise RuntimeError(_native.last_error())
    out = (assign, any4, sz)
    if pack_inner_k is not None:
        out += (packed,)
    if return_lut:
        out += (lut,)
    return out


def any4_linear_from_float(linear, group_size=128, w_inner_k=4, sample_weight=None, **kw):
    """torch.nn.Linear (weights on a CUDA device, bf16 / fp16) -> packed `Any4Linear`, ready for the GEMV kernels: the
    quantizer writes the tensor-core layout directly (no [n][k] int32 code matrix, no convert pass).  The analogue of
    quantize.py:880-900 (anyq_layer with a tinygemm pseudo=False module)."""
    from .modules import Any4Linear

    W = linear.weight.data
    n, k = W.shape
    _, any4, sz, packed, lut = anyq_quantize_tensor(W, q_group_size=group_size, sample_weight=sample_weight,
                                                    pack_inner_k=w_inner_k, return_lut=True, **kw)
    q = Any4Linear(k, n, bias=linear.bias is not None, device="meta", dtype=W.dtype, group_size=group_size,
                   w_inner_k=w_inner_k, per_row=True)
    q.weight = torch.nn.Parameter(packed, requires_grad=False)
    q.scales_and_zeros = torch.nn.Parameter(sz, requires_grad=False)
    q.lut = torch.nn.Parameter(lut, requires_grad=False)
    if linear.bias is not None:
        q.bias = torch.nn.Parameter(linear.bias.data.to(W.dtype).clone(), requires_grad=False)
    q.weight_reshaped = True
    return q
