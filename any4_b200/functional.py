"""Drop-in for `tinygemm_lib.functional` (reference: tinygemm_lib/functional.py:10-259).

Same 16 function names, positional order, defaults and return layouts; each call lowers to the
same `torch.ops.tinygemm.*` sequence the reference issues (convert -> GEMM -> convert for the
`f16TC` variants, one GEMM op for the `f16RM` variants), implemented by the B200 library.

Naming:  linear_y_<L>_x_<L>_W_<fmt>TC  -> y = x @ W^T, weight is the mma "B" operand (weightOnRight)
         linear_y_<L>_W_<fmt>TC_x_<L>  -> same product with the weight as the mma "A" operand
         <L> = f16TC (activations/outputs in tensor-core layout) or f16RM (row-major).
"""
import torch

from . import _native

_native.load_ops()
_ops = torch.ops.tinygemm

# weight innerKTiles accepted per API (functional.py:10-18 in the reference lists the any4 rows)
_VALID_W_INNER_K = {
    "linear_y_f16RM_x_f16RM_W_any4TC": (2, 4, 8),
    "linear_y_f16TC_x_f16TC_W_any4TC": (2, 4, 8),
    "linear_y_f16TC_W_any4TC_x_f16TC": (1, 2, 4),
    "linear_y_f16RM_W_any4TC_x_f16RM": (1, 2, 4),
}


def set_static_weights(flag=True):
    """B200 extension (not in the reference): promise that packed weights / LUTs / scales are not produced by the
    kernel launched right before a GEMV on the same stream, so the weight stream of each GEMV may start while the
    previous kernel is still finishing (include/tinygemm_b200.h: TG_OPT_STATIC_WEIGHTS).  True for a model whose
    layers were packed once with `reshape_weight()`; keep it off when calling the `reshape_weight=True` wrappers."""
    rc = _native.capi().tg_set_option(1, 1 if flag else 0)
    if rc != 0:
        raise RuntimeError(_native.last_error())


def valid_tinygemm_kernel_call(functional_api, w_inner_k):
    """True when (api, w_inner_k) is a supported any4 combination, else None (as the reference)."""
    if w_inner_k in _VALID_W_INNER_K.get(functional_api, ()):
        return True


# ---- weight packers: row-major codes / values -> tensor-core layout (no-op when pre-packed) ----
def _pack(w, layout, inner_k, do_pack):
    if not do_pack:
        return w
    return getattr(_ops, f"convert_matrix_to_m16n8k16_{layout}_layout")(w, inner_k)


def _tc_right(gemm, x, w2, x_inner_k, out_cols, *qargs):
    """activations -> A layout, GEMM with the weight on the right, output (A layout) -> row-major"""
    x2 = _ops.convert_matrix_to_m16n8k16_A_layout(x, x_inner_k)
    y2 = gemm(x2, w2, *qargs, True)
    return _ops.convert_matrix_from_m16n8k16_A_layout(y2, x.size(0), out_cols)


def _tc_left(gemm, x, w2, x_inner_k, out_cols, *qargs):
    """activations -> B layout, GEMM with the weight on the left, output (B layout) -> row-major"""
    x2 = _ops.convert_matrix_to_m16n8k16_B_layout(x, x_inner_k)
    y2 = gemm(w2, x2, *qargs, False)
    return _ops.convert_matrix_from_m16n8k16_B_layout(y2, x.size(0), out_cols)


# ------------------------------------------------------------------------------------------
# int4
# ------------------------------------------------------------------------------------------
def linear_y_f16TC_x_f16TC_W_int4TC(x, w_int32, w_scales_and_zeros, q_group, w_inner_k=4, x_inner_k=1,
                                    reshape_weight=True):
    w2 = _pack(w_int32, "Bint4", w_inner_k, reshape_weight)
    return _tc_right(_ops.tinygemm_y_f16TC_x_f16TC_w_int4TC, x, w2, x_inner_k, w_int32.size(0),
                     q_group, w_scales_and_zeros)


def linear_y_f16TC_W_int4TC_x_f16TC(x, w_int32, w_scales_and_zeros, q_group, w_inner_k=4, x_inner_k=1,
                                    reshape_weight=True):
    w2 = _pack(w_int32, "Aint4", w_inner_k, reshape_weight)
    return _tc_left(_ops.tinygemm_y_f16TC_x_f16TC_w_int4TC, x, w2, x_inner_k, w_int32.size(0),
                    q_group, w_scales_and_zeros)


def linear_y_f16RM_x_f16RM_W_int4TC(x, w_int32, w_scales_and_zeros, q_group, w_inner_k=4, reshape_weight=True):
    w2 = _pack(w_int32, "Bint4", w_inner_k, reshape_weight)
    return _ops.tinygemm_y_f16RM_x_f16RM_w_int4TC(x, w2, q_group, w_scales_and_zeros, True)


def linear_y_f16RM_W_int4TC_x_f16RM(x, w_int32, w_scales_and_zeros, q_group, w_inner_k=4, reshape_weight=True):
    w2 = _pack(w_int32, "Aint4", w_inner_k, reshape_weight)
    return _ops.tinygemm_y_f16RM_x_f16RM_w_int4TC(w2, x, q_group, w_scales_and_zeros, False)


# ------------------------------------------------------------------------------------------
# int8
# ------------------------------------------------------------------------------------------
def linear_y_f16TC_x_f16TC_W_int8TC(x, w_int32, w_scales_and_zeros, q_group, w_inner_k=4, reshape_weight=True):
    w2 = _pack(w_int32, "Bint8", w_inner_k, reshape_weight)
    return _tc_right(_ops.tinygemm_y_f16TC_x_f16TC_w_int8TC, x, w2, 1, w_int32.shape[0],
                     q_group, w_scales_and_zeros)


def linear_y_f16TC_W_int8TC_x_f16TC(x, w_int32, w_scales_and_zeros, q_group, w_inner_k=4, x_inner_k=1,
                                    reshape_weight=True):
    w2 = _pack(w_int32, "Aint8", w_inner_k, reshape_weight)
    # NOTE: the reference un-packs the result with n = x.shape[1] here (functional.py:113), which is
    # only right when in_features == out_features; the weight's row count is what is meant.
    return _tc_left(_ops.tinygemm_y_f16TC_x_f16TC_w_int8TC, x, w2, x_inner_k, w_int32.shape[0],
                    q_group, w_scales_and_zeros)


def linear_y_f16RM_x_f16RM_W_int8TC(x, w_int32, w_scales_and_zeros, q_group, w_inner_k=4, reshape_weight=True):
    w2 = _pack(w_int32, "Bint8", w_inner_k, reshape_weight)
    return _ops.tinygemm_y_f16RM_x_f16RM_w_int8TC(x, w2, q_group, w_scales_and_zeros, True)


def linear_y_f16RM_W_int8TC_x_f16RM(x, w_int32, w_scales_and_zeros, q_group, w_inner_k=4, reshape_weight=True):
    w2 = _pack(w_int32, "Aint8", w_inner_k, reshape_weight)
    return _ops.tinygemm_y_f16RM_x_f16RM_w_int8TC(w2, x, q_group, w_scales_and_zeros, False)


# ------------------------------------------------------------------------------------------
# any4 (LUT: [16] for one table per matrix - nf4/fp4/af4 - or [rows][16] per weight row)
# ------------------------------------------------------------------------------------------
def linear_y_f16TC_x_f16TC_W_any4TC(x, w_int32, w_lut, w_scales_and_zeros, q_group, w_inner_k=4, x_inner_k=1,
                                    reshape_weight=True):
    w2 = _pack(w_int32, "Bint4", w_inner_k, reshape_weight)
    return _tc_right(_ops.tinygemm_y_f16TC_x_f16TC_w_any4TC, x, w2, x_inner_k, w_int32.size(0),
                     q_group, w_scales_and_zeros, w_lut)


def linear_y_f16TC_W_any4TC_x_f16TC(x, w_int32, w_lut, w_scales_and_zeros, q_group, w_inner_k=4, x_inner_k=1,
                                    reshape_weight=True):
    w2 = _pack(w_int32, "Aint4", w_inner_k, reshape_weight)
    return _tc_left(_ops.tinygemm_y_f16TC_x_f16TC_w_any4TC, x, w2, x_inner_k, w_int32.size(0),
                    q_group, w_scales_and_zeros, w_lut)


def linear_y_f16RM_x_f16RM_W_any4TC(x, w_int32, w_lut, w_scales_and_zeros, q_group, w_inner_k=4,
                                    reshape_weight=True):
    w2 = _pack(w_int32, "Bint4", w_inner_k, reshape_weight)
    return _ops.tinygemm_y_f16RM_x_f16RM_w_any4TC(x, w2, q_group, w_scales_and_zeros, w_lut, True)


def linear_y_f16RM_W_any4TC_x_f16RM(x, w_int32, w_lut, w_scales_and_zeros, q_group, w_inner_k=4,
                                    reshape_weight=True):
    w2 = _pack(w_int32, "Aint4", w_inner_k, reshape_weight)
    return _ops.tinygemm_y_f16RM_x_f16RM_w_any4TC(w2, x, q_group, w_scales_and_zeros, w_lut, False)


# ------------------------------------------------------------------------------------------
# 16-bit weights
# ------------------------------------------------------------------------------------------
def linear_y_f16TC_x_f16TC_W_f16TC(x, w, w_inner_k=4, reshape_weight=True):
    w2 = _pack(w, "B", w_inner_k, reshape_weight)
    return _tc_right(_ops.tinygemm_y_f16TC_x_f16TC_w_f16TC, x, w2, 1, w.shape[0])


def linear_y_f16TC_W_f16TC_x_f16TC(x, w, x_inner_k=4, reshape_weight=True):
    w2 = _pack(w, "A", 1, reshape_weight)
    return _tc_left(_ops.tinygemm_y_f16TC_x_f16TC_w_f16TC, x, w2, x_inner_k, w.shape[0])


def linear_y_f16RM_x_f16RM_W_f16TC(x, w, w_inner_k=4, reshape_weight=True):
    w2 = _pack(w, "B", w_inner_k, reshape_weight)
    return _ops.tinygemm_y_f16RM_x_f16RM_w_f16TC(x, w2, True)


def linear_y_f16RM_W_f16TC_x_f16RM(x, w, w_inner_k=4, reshape_weight=True):
    w2 = _pack(w, "A", w_inner_k, reshape_weight)
    return _ops.tinygemm_y_f16RM_x_f16RM_w_f16TC(w2, x, False)
