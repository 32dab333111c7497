"""CPU: the C-ABI library loads, exports every symbol include/tinygemm_b200.h declares, validates
arguments before touching CUDA, and the torch op layer registers the reference's 19 schemas."""
import ctypes
import os
import re

import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def lib():
    from any4_b200 import _native, build

    build.build_all()
    return _native.capi()


def test_every_declared_symbol_is_exported(lib):
    from any4_b200 import _native

    hdr = open(os.path.join(ROOT, "include", "tinygemm_b200.h")).read()
    declared = set(re.findall(r"\b(tg_[a-z0-9_A-Z]+)\s*\(", hdr))
    assert declared == set(_native.CAPI_SYMBOLS)
    for name in declared:
        assert getattr(lib, name) is not None
    assert lib.tg_version().startswith(b"tinygemm_b200")


def test_argument_validation_without_gpu(lib):
    one = ctypes.c_void_p(256)
    assert lib.tg_convert_to_Bint4(one, one, 8, 96, 4, None) == -1      # k % 64 != 0
    assert b"multiple" in lib.tg_last_error()
    assert lib.tg_convert_to_Bint4(one, one, 8, 128, 3, None) == -1     # bad innerKTiles
    assert lib.tg_convert_to_Aint8(one, one, 8, 128, 4, None) == -1
    assert lib.tg_convert_to_B(one, one, 8, 128, 4, None) == -1
    # GEMM: k % 32, group, inner k, mx4 + fp16
    args = dict(y=one, x=one, w=one, sz=one, lut=one, e=one)
    def w4(rows_x=1, w_rows=64, k=256, group=128, ik=4, fmt=2, side=1, dt=0):
        return lib.tg_gemm_w4_rm(args["y"], args["x"], args["w"], args["sz"], args["lut"], args["e"],
                                 rows_x, w_rows, k, group, ik, fmt, side, dt, None)
    assert w4(k=200) == -1
    assert w4(group=48) == -1
    assert w4(ik=1) == -1            # B layout needs 2, 4, 8
    assert w4(ik=8, side=0) == -1    # A layout needs 1, 2, 4
    assert w4(w_rows=60) == -1
    assert w4(fmt=3, dt=1, group=32) == -3   # mx4 is bf16 only (TG_ERR_UNSUPPORTED)
    assert w4(rows_x=0) == 0         # empty activation: nothing to launch
    assert lib.tg_set_option(0, 1) == 0 and lib.tg_set_option(1, 0) == 0 and lib.tg_set_option(9, 1) == -1
    assert lib.tg_gemm_tc_workspace_bytes(16, 64, 256) >= 16 * 256 * 2 + 16 * 64 * 2
    # round-2 entry points
    assert lib.tg_gemm_w4_rm_workspace_bytes(1, 64, 256, 0) == 0            # one row: the A kernel itself
    assert lib.tg_gemm_w4_rm_workspace_bytes(8, 64, 256, 0) == 64 * 256 // 2  # several rows, A layout: the B repack
    assert lib.tg_gemm_w4_rm_workspace_bytes(8, 64, 256, 1) == 0            # B layout: nothing to repack
    assert lib.tg_repack_Aint4_to_Bint4(one, one, 24, 256, 4, 4, None) == -1  # rows not the padded A row count
    assert lib.tg_repack_Aint4_to_Bint4(one, one, 32, 96, 4, 4, None) == -1   # k % 64
    f = ctypes.c_float
    assert lib.tg_quantize_any4_rows(one, None, 8, 200, 128, 4, 300, f(1e-4), None, None, one, one, one, None, 0, None) == -1
    assert lib.tg_quantize_any4_rows(one, None, 8, 256, 48, 4, 300, f(1e-4), None, None, one, one, one, None, 0, None) == -1
    assert lib.tg_quantize_any4_rows(one, None, 12, 256, 128, 4, 300, f(1e-4), None, one, one, one, one, None, 0, None) == -1  # packed: n % 8
    peers = (ctypes.c_void_p * 2)(256, 512)
    assert lib.tg_gemm_w4_rm_exchange(one, peers, 0, 0, 2, 128, one, one, one, one, None, 1, 64, 256, 128, 4, 2, 0, None) == -1  # tag 0
    assert lib.tg_gemm_w4_rm_exchange(one, peers, 5, 1, 2, 128, one, one, one, one, None, 1, 64, 256, 128, 4, 2, 0, None) == -1  # bad rank
    assert lib.tg_gemm_w4_rm_exchange(one, peers, 0, 1, 2, 64, one, one, one, one, None, 1, 64, 256, 128, 4, 2, 0, None) == -1   # stride < full row
    assert lib.tg_gemm_w4_rm_exchange_silu_pairs(one, peers, 0, 1, 2, 64, one, one, one, one, None, 1, 72, 256, 128, 4, 2, 0,
                                                 None) == -1                     # stride 64 < 2 * 72 / 2 outputs
    assert lib.tg_gemm_w4_rm_hostio(one, one, None, one, one, one, None, 1, 64, 256, 128, 4, 2, 1, 0, None) == -1               # no staging buffer


EXPECTED_OPS = [
    "convert_matrix_to_m16n8k16_A_layout", "convert_matrix_to_m16n8k16_Aint4_layout",
    "convert_matrix_to_m16n8k16_Aint8_layout", "convert_matrix_from_m16n8k16_A_layout",
    "convert_matrix_to_m16n8k16_B_layout", "convert_matrix_to_m16n8k16_Bint4_layout",
    "convert_matrix_to_m16n8k16_Bint8_layout", "convert_matrix_from_m16n8k16_B_layout",
    "tinygemm_y_f16TC_x_f16TC_w_int4TC", "tinygemm_y_f16RM_x_f16RM_w_int4TC",
    "tinygemm_y_f16TC_x_f16TC_w_any4TC", "tinygemm_y_f16RM_x_f16RM_w_any4TC",
    "tinygemm_y_f16TC_x_f16TC_w_mx4TC", "tinygemm_y_f16RM_x_f16RM_w_mx4TC",
    "tinygemm_y_f16TC_x_f16TC_w_int8TC", "tinygemm_y_f16RM_x_f16RM_w_int8TC",
    "tinygemm_y_f16TC_x_f16TC_w_f16TC", "tinygemm_y_f16RM_x_f16RM_w_f16TC",
    "tinygemm_dequant_int4",
]


def test_op_schemas_match_reference(lib):
    import tinygemm  # noqa: F401

    for name in EXPECTED_OPS:
        assert hasattr(torch.ops.tinygemm, name), name
    s = str(torch.ops.tinygemm.tinygemm_y_f16RM_x_f16RM_w_any4TC.default._schema)
    assert s == ("tinygemm::tinygemm_y_f16RM_x_f16RM_w_any4TC(Tensor A, Tensor B, int qGroupSize, "
                 "Tensor qScaleAndZeros, Tensor int4DequantValues, bool weightOnRight) -> Tensor")
    ref_cpp = "/root/reference/tinygemm_lib/TinyGemm.cpp"
    if os.path.exists(ref_cpp):  # build container only: compare every schema string with the reference's
        src = open(ref_cpp).read()
        frag = src[src.index("TORCH_LIBRARY_FRAGMENT"):src.index("TORCH_LIBRARY(tinygemm")]
        defs = re.findall(r"m\.def\(\s*((?:\"[^\"]*\"\s*)+)\)", frag)
        assert len(defs) == 19
        for d in defs:
            schema = "".join(re.findall(r"\"([^\"]*)\"", d))
            name = schema.split("(")[0]
            ours = str(getattr(torch.ops.tinygemm, name).default._schema)
            assert ours == "tinygemm::" + schema, (ours, schema)


def test_ops_fail_loudly_on_cpu_tensors(lib):
    import tinygemm  # noqa: F401

    with pytest.raises(RuntimeError):
        torch.ops.tinygemm.convert_matrix_to_m16n8k16_Bint4_layout(torch.zeros(8, 64, dtype=torch.int32), 4)
    x = torch.zeros(1, 64, dtype=torch.bfloat16)
    w = torch.zeros(1, 1, 32, 2, dtype=torch.int32)
    sz = torch.zeros(1, 8, 2, dtype=torch.bfloat16)
    with pytest.raises(RuntimeError):
        torch.ops.tinygemm.tinygemm_y_f16RM_x_f16RM_w_int4TC(x, w, 64, sz, True)


def test_functional_and_module_surface():
    import inspect

    import tinygemm_lib.functional as F
    from any4_b200 import modules

    names = [n for n in dir(F) if n.startswith("linear_y_")]
    assert len(names) == 16
    assert F.valid_tinygemm_kernel_call("linear_y_f16RM_x_f16RM_W_any4TC", 4) is True
    assert F.valid_tinygemm_kernel_call("linear_y_f16RM_x_f16RM_W_any4TC", 1) is None
    assert F.valid_tinygemm_kernel_call("linear_y_f16RM_W_any4TC_x_f16RM", 1) is True
    ref_py = "/root/reference/tinygemm_lib/functional.py"
    if os.path.exists(ref_py):  # same names, parameter names and defaults as the reference
        import ast

        tree = ast.parse(open(ref_py).read())
        for fn in [n for n in tree.body if isinstance(n, ast.FunctionDef) and n.name.startswith("linear_y_")]:
            ours = inspect.signature(getattr(F, fn.name))
            ref_args = [a.arg for a in fn.args.args]
            assert list(ours.parameters) == ref_args, fn.name
            ref_defaults = [ast.literal_eval(d) for d in fn.args.defaults]
            our_defaults = [p.default for p in ours.parameters.values() if p.default is not inspect._empty]
            assert our_defaults == ref_defaults, fn.name
    m = modules.Any4Linear(256, 64, bias=True, dtype=torch.bfloat16, group_size=64)
    shapes = {n: tuple(p.shape) for n, p in m.named_parameters()}
    assert shapes == {"weight": (64, 256), "scales_and_zeros": (4, 64, 2), "lut": (64, 16), "bias": (64,)}
    assert m.weight.dtype == torch.int32 and m.kernel == "linear_y_f16RM_x_f16RM_W_any4TC" and m.w_inner_k == 4
    assert modules.Int4Linear(256, 64).kernel == "linear_y_f16RM_W_int4TC_x_f16RM"
    assert modules.Int8Linear(256, 64).w_inner_k == 2
    assert tuple(modules.Any4Linear(256, 64, per_row=False).lut.shape) == (16,)
