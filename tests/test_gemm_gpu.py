"""GPU parity: every GEMM op of the surface vs the CPU oracle (oracle/dequant.py) on seeded
inputs, through torch.ops.tinygemm -> C ABI -> CUDA kernels."""
import pytest
import torch

from oracle import cases as C
from oracle import dequant
from tests import _tol

pytestmark = pytest.mark.gpu

GEMM_CASES = C.gemm_cases()


def _dense_weight(c, inp):
    dt = C.DT[c["dt"]]
    n, g, fmt = c["n"], c["g"], c["fmt"]
    if fmt == "f16":
        return inp["w"]
    if fmt == "int4":
        return dequant.dequant_int4(inp["codes"], inp["sz"][:, :n], g, dt)
    if fmt == "int8":
        return dequant.dequant_int8(inp["codes"], inp["sz"][:, :n], g, dt)
    if fmt == "any4g":
        return dequant.dequant_lut(inp["codes"], inp["lut"], inp["sz"][:, :n], g, dt)
    if fmt == "any4r":
        return dequant.dequant_lut(inp["codes"], inp["lut"][:n], inp["sz"][:, :n], g, dt)
    return dequant.dequant_mx4(inp["codes"], inp["exps"][:n], g, dt)


@pytest.mark.parametrize("case", GEMM_CASES, ids=[C.case_id(c) for c in GEMM_CASES])
def test_gemm_matches_oracle(case, cuda_device):
    import tinygemm  # noqa: F401

    inp = C.make_inputs(case)
    got = C.run_ops(case, inp, cuda_device)["y"]
    dt = C.DT[case["dt"]]
    w = _dense_weight(case, inp)
    x = inp["x"]
    y64 = dequant.gemm_f64(x, w)
    absdot = x.double().abs() @ w.double().abs().t()
    assert got.shape == y64.shape and got.dtype == dt
    assert torch.isfinite(got.float()).all()
    nbad, worst = _tol.check_faithful(got, y64, absdot, dt)
    assert nbad == 0, f"{nbad} elements outside the faithful-rounding bound (worst {worst:.2f}x)"
    ref = dequant.gemm(x, w)
    assert _tol.frob_rel(got, ref) <= _tol.FROB_REL
    ulp = dequant.ulp_distance(got, ref)
    assert (ulp == 0).double().mean().item() >= 0.95


def _int8_case(i, **kw):
    return dict(kind="gemm", seed=9100 + i, fmt="int8", api="RM", x_ik=1, **kw)


# The int8 ring kernel (csrc/gemv_generic.cu gemm_w8_ring_kernel): several chunks per row tile, a partial last chunk,
# every inner-k / group size, two passes of activation rows, more row tiles than resident CTAs (ring wrap-around and
# the double-buffered partial sums), and a B-layout row count that is not a multiple of 16 (stream-kernel fallback).
INT8_RING_CASES = [_int8_case(i, **kw) for i, kw in enumerate([
    dict(dt="bf16", side="right", m=1, n=64, k=4096, g=128, ik=4),
    dict(dt="fp16", side="right", m=8, n=48, k=2112, g=64, ik=4),
    dict(dt="bf16", side="right", m=9, n=32, k=1024, g=32, ik=4),
    dict(dt="bf16", side="right", m=16, n=40, k=1024, g=256, ik=2),
    dict(dt="fp16", side="right", m=2, n=16, k=1056, g=32, ik=1),
    dict(dt="bf16", side="right", m=5, n=80, k=3072, g=256, ik=2),
    dict(dt="bf16", side="right", m=2, n=16 * (2 * 148 + 5), k=1024, g=128, ik=4),
    dict(dt="bf16", side="right", m=3, n=32, k=1056, g=32, ik=2),
    dict(dt="bf16", side="left", m=3, n=48, k=3072, g=128, ik=1),
    dict(dt="fp16", side="left", m=8, n=16, k=1184, g=32, ik=2),
    dict(dt="bf16", side="left", m=12, n=32, k=2048, g=64, ik=2),
    dict(dt="bf16", side="left", m=1, n=64, k=2304, g=256, ik=1),
    dict(dt="fp16", side="left", m=4, n=16 * (2 * 148 + 3), k=512, g=128, ik=2),
])]


@pytest.mark.parametrize("case", INT8_RING_CASES, ids=[C.case_id(c) for c in INT8_RING_CASES])
def test_int8_ring_kernel_matches_oracle(case, cuda_device):
    test_gemm_matches_oracle(case, cuda_device)


# the same kernel with 16-bit weights (512 k per chunk, no group words)
F16_RING_CASES = [dict(kind="gemm", seed=9200 + i, fmt="f16", api="RM", x_ik=1, g=0, **kw) for i, kw in enumerate([
    dict(dt="bf16", side="right", m=1, n=64, k=4096, ik=2),
    dict(dt="fp16", side="right", m=8, n=48, k=2112, ik=2),
    dict(dt="bf16", side="right", m=9, n=32, k=1056, ik=1),
    dict(dt="bf16", side="right", m=16, n=40, k=1024, ik=2),           # 40 rows: stream-kernel fallback
    dict(dt="bf16", side="right", m=2, n=16 * (2 * 148 + 5), k=512, ik=2),
    dict(dt="bf16", side="left", m=3, n=48, k=3072, ik=1),
    dict(dt="fp16", side="left", m=12, n=32, k=1184, ik=1),
    dict(dt="bf16", side="left", m=4, n=16 * (2 * 148 + 3), k=256, ik=1),
])]


@pytest.mark.parametrize("case", F16_RING_CASES, ids=[C.case_id(c) for c in F16_RING_CASES])
def test_f16_ring_kernel_matches_oracle(case, cuda_device):
    test_gemm_matches_oracle(case, cuda_device)


@pytest.mark.parametrize("dt", ["bf16", "fp16"])
@pytest.mark.parametrize("fmt", ["int4", "any4g", "mx4", "int8"])
@pytest.mark.parametrize("side", ["right", "left"])
def test_identity_weight_is_exact(fmt, side, dt, cuda_device):
    """The reference's strongest pin (tests/tinygemm/test_tinygemm_{int4,any4,mx4,int8}.py
    test_identity_mul): W = I quantised with the test-side quantizer gives y == x bit-exactly in
    bf16 (fp16 int4/int8 is knowingly inexact there: 15 * fp16(1/15) != 1, checked to 1e-3)."""
    import tinygemm  # noqa: F401
    from any4_b200 import utils as host

    if fmt == "mx4" and dt == "fp16":
        pytest.skip("mx4 is bf16 only")
    ops = torch.ops.tinygemm
    tdt = C.DT[dt]
    k = 256
    right = side == "right"
    gen = torch.Generator().manual_seed(7)
    x = torch.randn(5, k, generator=gen).to(tdt).to(cuda_device)
    w = torch.eye(k, dtype=tdt, device=cuda_device)
    g = 32
    if fmt == "mx4":
        codes, e = host.quantize_mx4(w, g)
        w2 = (ops.convert_matrix_to_m16n8k16_Bint4_layout(codes, 4) if right
              else ops.convert_matrix_to_m16n8k16_Aint4_layout(codes, 4))
        A, B = (x, w2) if right else (w2, x)
        y = ops.tinygemm_y_f16RM_x_f16RM_w_mx4TC(A, B, g, e, right)
    elif fmt == "int8":
        codes, sz = host.group_quantize_tensor(w, 8, g)
        w2 = (ops.convert_matrix_to_m16n8k16_Bint8_layout(codes, 2) if right
              else ops.convert_matrix_to_m16n8k16_Aint8_layout(codes, 2))
        A, B = (x, w2) if right else (w2, x)
        y = ops.tinygemm_y_f16RM_x_f16RM_w_int8TC(A, B, g, sz, right)
    else:
        codes, sz = host.group_quantize_tensor(w, 4, g)
        w2 = (ops.convert_matrix_to_m16n8k16_Bint4_layout(codes, 4) if right
              else ops.convert_matrix_to_m16n8k16_Aint4_layout(codes, 4))
        A, B = (x, w2) if right else (w2, x)
        if fmt == "int4":
            y = ops.tinygemm_y_f16RM_x_f16RM_w_int4TC(A, B, g, sz, right)
        else:
            lut = (torch.arange(16, dtype=torch.float32) - 8).to(tdt).to(cuda_device)
            y = ops.tinygemm_y_f16RM_x_f16RM_w_any4TC(A, B, g, sz, lut, right)
    if dt == "bf16" or fmt == "mx4":
        assert torch.equal(y, x)
    else:
        assert (y.float() - x.float()).abs().max().item() < 5e-3


def test_mx4_nan_exponent(cuda_device):
    """reference: tests/tinygemm/test_tinygemm_mx4.py:443-506 - e = 254 stays finite, e = 255 gives NaN"""
    import tinygemm  # noqa: F401
    from any4_b200 import utils as host

    ops = torch.ops.tinygemm
    k = 32
    x = torch.randn(5, k, dtype=torch.bfloat16, device=cuda_device)
    w = torch.eye(k, dtype=torch.bfloat16, device=cuda_device)
    codes, e = host.quantize_mx4(w, 32)
    w2 = ops.convert_matrix_to_m16n8k16_Bint4_layout(codes, 2)
    y = ops.tinygemm_y_f16RM_x_f16RM_w_mx4TC(x, w2, 32, e, True)
    assert torch.equal(y, x)
    e[0][0] = 254
    y = ops.tinygemm_y_f16RM_x_f16RM_w_mx4TC(x, w2, 32, e, True)
    assert not torch.isnan(y).any()
    e[0][0] = 255
    y = ops.tinygemm_y_f16RM_x_f16RM_w_mx4TC(x, w2, 32, e, True)
    assert torch.isnan(y).any()


def test_lut_is_honoured(cuda_device):
    """reference: tests/tinygemm/test_tinygemm_any4.py:17-26 - negating LUT and scales together
    must leave the product unchanged, proving the LUT (not a built-in int4 table) is used."""
    import tinygemm  # noqa: F401
    from any4_b200 import utils as host

    ops = torch.ops.tinygemm
    dev = cuda_device
    gen = torch.Generator().manual_seed(3)
    x = torch.randn(3, 256, generator=gen).bfloat16().to(dev)
    w = torch.randn(64, 256, generator=gen).bfloat16().to(dev)
    codes, sz = host.group_quantize_tensor(w, 4, 64)
    w2 = ops.convert_matrix_to_m16n8k16_Bint4_layout(codes, 4)
    lut = (torch.arange(16, dtype=torch.float32) - 8).bfloat16().to(dev)
    y_int4 = ops.tinygemm_y_f16RM_x_f16RM_w_int4TC(x, w2, 64, sz, True)
    y_any4 = ops.tinygemm_y_f16RM_x_f16RM_w_any4TC(x, w2, 64, sz, lut, True)
    assert torch.equal(y_int4, y_any4)
    sz_neg = sz.clone()
    sz_neg[:, :, 0] *= -1
    y_neg = ops.tinygemm_y_f16RM_x_f16RM_w_any4TC(x, w2, 64, sz_neg, -lut, True)
    assert torch.equal(y_neg, y_any4)


def test_large_shape_linearity_and_rows(cuda_device):
    """Full-size property checks at the BASELINE shape (4096 x 4096, g = 128), where the oracle is
    too slow for every element: (a) 64 sampled output columns against the float64 oracle,
    (b) row m of a batched call equals the m = 1 call on that row, (c) exact zero for x = 0."""
    import tinygemm  # noqa: F401

    ops = torch.ops.tinygemm
    dev = cuda_device
    n = k = 4096
    g = 128
    gen = torch.Generator().manual_seed(11)
    codes = torch.randint(0, 16, (n, k), generator=gen, dtype=torch.int32)
    lut = ((torch.rand(n, 16, generator=gen) * 15).sort(1).values.bfloat16() - 8)
    sz = torch.stack([torch.rand(k // g, n, generator=gen) * 0.01 + 0.001,
                      torch.randn(k // g, n, generator=gen) * 0.01], dim=2).bfloat16()
    x = torch.randn(4, k, generator=gen).bfloat16()
    w2 = ops.convert_matrix_to_m16n8k16_Bint4_layout(codes.to(dev), 4)
    lut_d, sz_d, x_d = lut.to(dev), sz.to(dev), x.to(dev)
    y1 = ops.tinygemm_y_f16RM_x_f16RM_w_any4TC(x_d[:1].contiguous(), w2, g, sz_d, lut_d, True)
    y4 = ops.tinygemm_y_f16RM_x_f16RM_w_any4TC(x_d, w2, g, sz_d, lut_d, True)
    rows = torch.randperm(n, generator=gen)[:64]
    wd = dequant.dequant_lut(codes[rows], lut[rows], sz[:, rows], g, torch.bfloat16)
    y64 = dequant.gemm_f64(x, wd)
    absdot = x.double().abs() @ wd.double().abs().t()
    nbad, worst = _tol.check_faithful(y4.cpu()[:, rows], y64, absdot, torch.bfloat16)
    assert nbad == 0, (nbad, worst)
    nbad, worst = _tol.check_faithful(y1.cpu()[:, rows], y64[:1], absdot[:1], torch.bfloat16)
    assert nbad == 0, (nbad, worst)
    # m = 1 and m = 4 paths use different mma operand roles: same value up to fp32 summation order
    assert _tol.frob_rel(y4[:1].cpu(), y1.cpu()) <= _tol.FROB_REL
    y0 = ops.tinygemm_y_f16RM_x_f16RM_w_any4TC(torch.zeros_like(x_d[:1]), w2, g, sz_d, lut_d, True)
    assert (y0 == 0).all()


@pytest.mark.parametrize("fmt,g", [("any4r", 32), ("mx4", 32), ("any4r", 64)])
def test_mma_sync_kernel_k_forced_split_with_many_row_blocks(fmt, g):
    """k beyond the mma.sync kernel's staging area forces a cluster split-k WITH more row blocks than
    clusters fit (n = 4096, k = 8192, group 32 / 64: 128 row blocks, splits >= 2).  Force that kernel
    (TG_OPT_W4_KERNEL = 2) and compare with the tcgen05 kernel, which is pinned to the reference at this size."""
    import tinygemm  # noqa: F401
    from any4_b200 import _native

    ops, lib = torch.ops.tinygemm, _native.capi()
    dev = torch.device("cuda:0")
    n, k = 4096, 8192
    gen = torch.Generator(device=dev).manual_seed(g)
    x = torch.randn(2, k, device=dev, generator=gen).bfloat16()
    w = torch.randint(-2**31, 2**31 - 1, (n // 8, k // 64, 32, 2), generator=gen, device=dev, dtype=torch.int64).to(torch.int32)
    if fmt == "mx4":
        exps = torch.randint(120, 130, (n, k // 32), generator=gen, device=dev, dtype=torch.int32).to(torch.uint8)
        run = lambda: ops.tinygemm_y_f16RM_x_f16RM_w_mx4TC(x, w, 32, exps, True)
    else:
        lut = (torch.rand(n, 16, device=dev, generator=gen) * 15).sort(1).values.bfloat16() - 8
        sz = torch.stack([torch.rand(k // g, n, generator=gen, device=dev) * 0.01 + 0.001,
                          torch.randn(k // g, n, generator=gen, device=dev) * 0.01], 2).bfloat16().contiguous()
        run = lambda: ops.tinygemm_y_f16RM_x_f16RM_w_any4TC(x, w, g, sz, lut, True)
    try:
        assert lib.tg_set_option(2, 1) == 0
        ref = run()
        assert lib.tg_set_option(2, 2) == 0
        got = run()
    finally:
        lib.tg_set_option(2, 0)
    assert ((got.float() - ref.float()).abs() <= 2.0 ** -7 * ref.float().abs() + 2.0 ** -9 * ref.float().abs().max()).all()
    assert (got == ref).float().mean() > 0.97


@pytest.mark.parametrize("fmt", ["any4r", "int4", "mx4"])
@pytest.mark.parametrize("m", [1, 3, 4])
def test_tcgen05_kernel_agrees_with_mma_sync_kernel(fmt, m):
    """B-layout 4-bit GEMM: the tcgen05 / TMEM kernel (TG_OPT_W4_KERNEL = 1) and the lane-per-row mma.sync kernel
    (= 2) dequantise identically and differ only in the order of the fp32 partial sums."""
    import tinygemm  # noqa: F401
    from any4_b200 import _native
    from any4_b200 import utils as U

    ops, lib = torch.ops.tinygemm, _native.capi()
    dev = torch.device("cuda:0")
    n, k, g = 264, 1024, 128
    gen = torch.Generator(device=dev).manual_seed(m + len(fmt))
    x = torch.randn(m, k, device=dev, generator=gen).bfloat16()
    wf = torch.randn(n, k, device=dev, generator=gen).bfloat16()
    if fmt == "mx4":
        codes, exps = U.quantize_mx4(wf, 32)
        w = ops.convert_matrix_to_m16n8k16_Bint4_layout(codes.to(torch.int32), 4)
        run = lambda: ops.tinygemm_y_f16RM_x_f16RM_w_mx4TC(x, w, 32, exps, True)
    else:
        codes, sz = U.group_quantize_tensor(wf, 4, g)
        w = ops.convert_matrix_to_m16n8k16_Bint4_layout(codes, 4)
        if fmt == "int4":
            run = lambda: ops.tinygemm_y_f16RM_x_f16RM_w_int4TC(x, w, g, sz, True)
        else:
            lut = (torch.rand(n, 16, device=dev, generator=gen) * 15).sort(1).values.bfloat16() - 8
            run = lambda: ops.tinygemm_y_f16RM_x_f16RM_w_any4TC(x, w, g, sz, lut, True)
    try:
        assert lib.tg_set_option(2, 2) == 0          # mma.sync kernel
        ref = run()
        assert lib.tg_set_option(2, 1) == 0          # tcgen05 kernel
        got = run()
    finally:
        lib.tg_set_option(2, 0)
    assert got.shape == ref.shape
    assert ((got.float() - ref.float()).abs() <= 2.0 ** -7 * ref.float().abs() + 1e-3).all()
    assert (got == ref).float().mean() > 0.9


@pytest.mark.parametrize("m", [1, 3])
def test_graph_replay_with_new_activations(m):
    """A captured launch is replayed with the launch tag it was captured with: the stream-K fix-up (row blocks shared by
    several CTAs, partial sums exchanged through tagged workspace words) must never take a PREVIOUS replay's partial
    for its own.  Replay a graph on changing activations and compare every replay with a plain launch."""
    import tinygemm  # noqa: F401
    from any4_b200 import _native
    from bench import synth_layer

    lib = _native.capi()
    dev = torch.device("cuda:0")
    n, k = 2048, 8192                                   # 64 row blocks x 8 stages over 148 CTAs: every row block is split
    w, lut, sz = synth_layer(n, k, 3, dev)
    op = torch.ops.tinygemm.tinygemm_y_f16RM_x_f16RM_w_any4TC
    gen = torch.Generator(device=dev).manual_seed(11 + m)
    x = torch.randn(m, k, device=dev, generator=gen).bfloat16()
    try:
        assert lib.tg_set_option(2, 1) == 0             # the tcgen05 kernel
        for _ in range(2):
            op(x, w, 128, sz, lut, True)
        torch.cuda.synchronize()
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g):
            ys = [op(x, w, 128, sz, lut, True) for _ in range(3)]   # back to back: programmatic dependent launches
        for rep in range(4):
            x.copy_(torch.randn(m, k, device=dev, generator=gen).bfloat16() * (rep + 1))
            g.replay()
            torch.cuda.synchronize()
            want = op(x, w, 128, sz, lut, True)
            for y in ys:
                assert torch.equal(y, want), f"replay {rep}"
    finally:
        lib.tg_set_option(2, 0)


@pytest.mark.parametrize("api", ["linear_y_f16RM_W_int4TC_x_f16RM", "linear_y_f16RM_W_int8TC_x_f16RM"])
def test_reference_general_mul_case_against_fp32(api):
    """The reference's `test_general_mul` cases that fail on a B200 for the reference's OWN kernels as well as for ours
    (profiles/r2/reference_testsuite_*.txt; weight in {0, 1} on the left, 3 activation rows, k = 2048): they compare
    with a bf16 cuBLAS product whose error alone is at the 0.1 threshold.  Against the fp32 product of the same
    operands the kernels' average error is two orders of magnitude below it."""
    import tinygemm_lib.functional as TF
    from tinygemm_lib.utils import group_quantize_tensor

    dev = torch.device("cuda:0")
    gen = torch.Generator(device=dev).manual_seed(3)
    for (m, n, k, ik, g) in [(32, 3, 2048, 2, 64), (48, 3, 2048, 2, 256), (32, 3, 2048, 2, 128)]:
        w = torch.randint(0, 2, (m, k), device=dev, generator=gen).bfloat16()
        x = torch.randn(n, k, device=dev, generator=gen).bfloat16()
        codes, sz = group_quantize_tensor(w, n_bit=8 if "int8" in api else 4, q_group_size=g)
        y = getattr(TF, api)(x, codes, sz, g, ik)
        exact = (w.float() @ x.float().t()).t().bfloat16()  # the fp32 product, rounded once (weights 0 / 1 dequantise exactly)
        avg_err = float((exact.float() - y[:, :m].float()).abs().sum() / (m * n))
        assert avg_err < 1e-2, (api, m, n, k, avg_err)


@pytest.mark.parametrize("side", ["left", "right"])
def test_very_long_k_small_group(side):
    """k = 36864 at group 32 is more than the lane-per-row kernels can stage per cluster (32768); the
    reference accepts it.  Left: falls back to the per-k-tile kernel; right: the tcgen05 kernel takes any k."""
    import tinygemm_lib.functional as TF
    from tinygemm_lib.utils import group_quantize_tensor

    dev = torch.device("cuda:0")
    gen = torch.Generator().manual_seed(17)
    n, k, g, m = 32, 36864, 32, 2
    w = (torch.randn(n, k, generator=gen) * 0.1).bfloat16()
    x = torch.randn(m, k, generator=gen).bfloat16()
    codes, sz = group_quantize_tensor(w, n_bit=4, q_group_size=g)
    wd = dequant.dequant_int4(codes, sz, g, torch.bfloat16)
    want = dequant.gemm(x, wd)
    fn = TF.linear_y_f16RM_W_int4TC_x_f16RM if side == "left" else TF.linear_y_f16RM_x_f16RM_W_int4TC
    y = fn(x.to(dev), codes.to(dev), sz.to(dev), g, 4)
    assert _tol.frob_rel(y[:, :n].cpu(), want) <= _tol.FROB_REL
