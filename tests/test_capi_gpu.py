"""GPU: drive the C ABI directly through ctypes (no torch ops in between) with torch tensors as
plain device buffers - the binding INTEGRATION.md documents."""
import ctypes

import pytest
import torch

from oracle import cases as C
from oracle import dequant, layouts
from tests import _tol

pytestmark = pytest.mark.gpu


def P(t):
    return ctypes.c_void_p(t.data_ptr())


def test_capi_any4_gemv_and_convert(cuda_device):
    from any4_b200 import _native

    lib = _native.capi()
    dev = cuda_device
    case = dict(kind="gemm", fmt="any4r", dt="bf16", side="right", api="RM", m=1, n=96, k=640, g=128, ik=4, x_ik=1, seed=77)
    inp = C.make_inputs(case)
    n, k, g = case["n"], case["k"], case["g"]
    codes = inp["codes"].to(dev)
    packed = torch.empty(n // 8, k // 64, 32, 2, dtype=torch.int32, device=dev)
    st = ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)
    assert lib.tg_convert_to_Bint4(P(codes), P(packed), n, k, 4, st) == 0, lib.tg_last_error()
    assert torch.equal(packed.cpu(), torch.from_numpy(layouts.to_Bint4(inp["codes"].numpy(), 4)))
    x, sz, lut = inp["x"].to(dev), inp["sz"].to(dev), inp["lut"].to(dev)
    y = torch.empty(1, n, dtype=torch.bfloat16, device=dev)
    lib.tg_reset_launch_count()
    rc = lib.tg_gemm_w4_rm(P(y), P(x), P(packed), P(sz), P(lut), None, 1, n, k, g, 4, 2, 1, 0, st)
    assert rc == 0, lib.tg_last_error()
    assert lib.tg_launch_count() == 1
    w = dequant.dequant_lut(inp["codes"], inp["lut"], inp["sz"], g, torch.bfloat16)
    y64 = dequant.gemm_f64(inp["x"], w)
    absdot = inp["x"].double().abs() @ w.double().abs().t()
    nbad, worst = _tol.check_faithful(y.cpu(), y64, absdot, torch.bfloat16)
    assert nbad == 0, (nbad, worst)
    # error reporting
    assert lib.tg_gemm_w4_rm(P(y), P(x), P(packed), P(sz), P(lut), None, 1, n, k, 48, 4, 2, 1, 0, st) == -1
    assert b"qGroupSize" in lib.tg_last_error()


def test_capi_is_stream_ordered_and_graph_capturable(cuda_device):
    """The GEMV must be capture-safe (no sync, no allocation): SURVEY 7 'hard parts'."""
    import tinygemm  # noqa: F401

    dev = cuda_device
    case = dict(kind="gemm", fmt="any4r", dt="bf16", side="right", api="RM", m=1, n=256, k=1024, g=128, ik=4, x_ik=1, seed=5)
    inp = C.make_inputs(case)
    ops = torch.ops.tinygemm
    w2 = ops.convert_matrix_to_m16n8k16_Bint4_layout(inp["codes"].to(dev), 4)
    x, sz, lut = inp["x"].to(dev), inp["sz"].to(dev), inp["lut"].to(dev)
    eager = ops.tinygemm_y_f16RM_x_f16RM_w_any4TC(x, w2, 128, sz, lut, True)
    s = torch.cuda.Stream()
    s.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(s):
        for _ in range(2):
            ops.tinygemm_y_f16RM_x_f16RM_w_any4TC(x, w2, 128, sz, lut, True)
    torch.cuda.current_stream().wait_stream(s)
    graph = torch.cuda.CUDAGraph()
    with torch.cuda.graph(graph):
        y_g = ops.tinygemm_y_f16RM_x_f16RM_w_any4TC(x, w2, 128, sz, lut, True)
    x.copy_(inp["x"].to(dev) * 2)
    graph.replay()
    torch.cuda.synchronize()
    assert torch.equal(y_g, ops.tinygemm_y_f16RM_x_f16RM_w_any4TC(x, w2, 128, sz, lut, True))
    assert not torch.equal(y_g, eager)


def test_pdl_and_static_weights_keep_stream_order(cuda_device):
    """TG_OPT_PDL / TG_OPT_STATIC_WEIGHTS must not change results: a chain of GEMVs whose activations are produced
    by the previous kernel (the decode pattern) gives the same bits with every option combination."""
    import tinygemm  # noqa: F401
    from any4_b200 import _native

    lib = _native.capi()
    dev = cuda_device
    ops = torch.ops.tinygemm
    n = k = 1024
    layers = []
    for i in range(6):
        case = dict(kind="gemm", fmt="any4r", dt="bf16", side="right", api="RM", m=1, n=n, k=k, g=128, ik=4, x_ik=1, seed=900 + i)
        inp = C.make_inputs(case)
        layers.append((ops.convert_matrix_to_m16n8k16_Bint4_layout(inp["codes"].to(dev), 4), inp["sz"].to(dev), inp["lut"].to(dev)))
    x0 = torch.randn(1, k, generator=torch.Generator().manual_seed(1)).bfloat16().to(dev)

    def chain():
        x = x0
        for w, sz, lut in layers:
            y = ops.tinygemm_y_f16RM_x_f16RM_w_any4TC(x, w, 128, sz, lut, True)
            x = torch.tanh(y)  # a non-tinygemm kernel in between, and the next GEMV reads its output
            x = ops.tinygemm_y_f16RM_x_f16RM_w_any4TC(x, w, 128, sz, lut, True)  # GEMV straight after GEMV
            x = torch.tanh(x)
        return x

    results = []
    try:
        for pdl, static in ((0, 0), (1, 0), (1, 1)):
            assert lib.tg_set_option(0, pdl) == 0 and lib.tg_set_option(1, static) == 0
            for _ in range(3):
                results.append(chain().clone())
            torch.cuda.synchronize()
    finally:
        lib.tg_set_option(0, 1)
        lib.tg_set_option(1, 0)
    for r in results[1:]:
        assert torch.equal(r, results[0])
    assert lib.tg_set_option(7, 1) == -1
