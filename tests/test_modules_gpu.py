"""GPU: the module layer (reference: modules.py, tests/test_anyq.py:146-194, tests/test_intq.py)."""
import pytest
import torch

from oracle import dequant
from tests import _tol

pytestmark = pytest.mark.gpu


def _fill_any4(m, gen, per_row=True):
    n, k, g = m.out_features, m.in_features, m.group_size
    m.weight.data = torch.randint(0, 16, (n, k), generator=gen, dtype=torch.int32).to(m.weight.device)
    lut = ((torch.rand(n if per_row else 1, 16, generator=gen) * 15).sort(1).values.bfloat16() - 8)
    m.lut.data = (lut if per_row else lut[0]).contiguous().to(m.weight.device)
    sz = torch.stack([torch.rand(k // g, n, generator=gen) * 0.01 + 0.001, torch.randn(k // g, n, generator=gen) * 0.01], 2)
    m.scales_and_zeros.data = sz.bfloat16().to(m.weight.device)
    if m.bias is not None:
        m.bias.data = torch.randn(n, generator=gen).bfloat16().to(m.weight.device)


@pytest.mark.parametrize("kernel,ik", [("linear_y_f16RM_x_f16RM_W_any4TC", 4), ("linear_y_f16RM_x_f16RM_W_any4TC", 8),
                                       ("linear_y_f16RM_W_any4TC_x_f16RM", 2)])
@pytest.mark.parametrize("per_row", [True, False])
def test_any4_linear_forward(kernel, ik, per_row, cuda_device):
    from any4_b200.modules import Any4Linear

    gen = torch.Generator().manual_seed(1)
    m = Any4Linear(512, 128, bias=True, device=cuda_device, dtype=torch.bfloat16, group_size=128, kernel=kernel,
                   w_inner_k=ik, per_row=per_row)
    _fill_any4(m, gen, per_row)
    codes = m.weight.data.cpu()
    x = torch.randn(2, 3, 512, generator=gen).bfloat16()
    y_unpacked = m(x.to(cuda_device))        # reshape_weight=True path: packs on the fly
    m.reshape_weight(ik)
    assert m.weight_reshaped and m.weight.dim() == 4
    y = m(x.to(cuda_device))
    assert y.shape == (2, 3, 128) and torch.equal(y, y_unpacked)
    w = dequant.dequant_lut(codes, m.lut.data.cpu(), m.scales_and_zeros.data.cpu(), 128, torch.bfloat16)
    ref = (dequant.gemm(x.view(-1, 512), w).float() + m.bias.data.cpu().float()).bfloat16().view(2, 3, 128)
    assert _tol.frob_rel(y.cpu(), ref) <= 2e-3
    # state_dict round trip keeps the packed weight
    m2 = Any4Linear(512, 128, bias=True, device=cuda_device, dtype=torch.bfloat16, group_size=128, kernel=kernel,
                    w_inner_k=ik, per_row=per_row)
    m2.weight.data = torch.empty_like(m.weight.data)
    m2.load_state_dict(m.state_dict())
    m2.weight_reshaped = True
    assert torch.equal(m2(x.to(cuda_device)), y)


def test_int4_and_int8_linear(cuda_device):
    from any4_b200 import utils as host
    from any4_b200.modules import Int4Linear, Int8Linear

    gen = torch.Generator().manual_seed(2)
    w = torch.randn(64, 256, generator=gen).bfloat16()
    x = torch.randn(5, 256, generator=gen).bfloat16()
    # (the TC-layout kernels un-pack y with w.size(0), which is only right for un-packed weights - a quirk
    #  shared with the reference, functional.py:33-35 - so the module test sticks to the RM kernels)
    for cls, bits, kernels in ((Int4Linear, 4, ["linear_y_f16RM_W_int4TC_x_f16RM", "linear_y_f16RM_x_f16RM_W_int4TC"]),
                               (Int8Linear, 8, ["linear_y_f16RM_W_int8TC_x_f16RM", "linear_y_f16RM_x_f16RM_W_int8TC"])):
        codes, sz = host.group_quantize_tensor(w, bits, 64)
        wd = (dequant.dequant_int4 if bits == 4 else dequant.dequant_int8)(codes, sz, 64, torch.bfloat16)
        ref = dequant.gemm(x, wd)
        for kern in kernels:
            m = cls(256, 64, bias=False, device=cuda_device, dtype=torch.bfloat16, group_size=64, kernel=kern)
            m.weight.data, m.scales_and_zeros.data = codes.to(cuda_device), sz.to(cuda_device)
            m.reshape_weight(m.w_inner_k)
            y = m(x.to(cuda_device))
            assert y.shape == (5, 64)
            assert _tol.frob_rel(y.cpu(), ref) <= _tol.FROB_REL, kern
    with pytest.raises(ValueError):
        Int4Linear(256, 64, kernel="nope", device=cuda_device, dtype=torch.bfloat16).reshape_weight()


def test_row_sharded_linear_single_process(cuda_device):
    """world = 2 emulated in one process: the two shards' zero-padded outputs sum to the full output."""
    from any4_b200.modules import Any4Linear, RowShardedLinear

    gen = torch.Generator().manual_seed(3)
    m = Any4Linear(512, 256, bias=True, device=cuda_device, dtype=torch.bfloat16, group_size=128)
    _fill_any4(m, gen)
    m.reshape_weight(4)
    x = torch.randn(3, 512, generator=gen).bfloat16().to(cuda_device)
    y = m(x)
    parts = []
    for r in range(2):
        sh = RowShardedLinear(m, r, 2)
        sh.world = 1  # no process group here: take the local zero-padded buffer as is
        bias, sh.bias = sh.bias, None
        parts.append(sh(x))
    total = parts[0] + parts[1] + m.bias
    assert torch.equal(total, y)


@pytest.mark.parametrize("cls_name", ["NF4Linear", "FP4Linear"])
def test_fixed_table_linear(cls_name, cuda_device):
    """NF4 / FP4 module shells (the reference's TODO, modules.py:10): quantize a float weight, run through the any4
    kernel with the fixed global table, compare with the CPU dequantisation of the stored codes / scales."""
    import any4_b200.modules as M

    gen = torch.Generator().manual_seed(4)
    w = torch.randn(128, 512, generator=gen) * 0.05
    x = torch.randn(3, 512, generator=gen).bfloat16()
    lin = getattr(M, cls_name)(512, 128, bias=False, device=cuda_device, dtype=torch.bfloat16, group_size=64)
    lin.quantize_weight(w.to(cuda_device))
    assert lin.weight_reshaped and lin.weight.dim() == 4
    y = lin(x.to(cuda_device))
    # the same quantisation on the CPU, dequantised by the oracle
    ref_lin = getattr(M, cls_name)(512, 128, bias=False, device="cpu", dtype=torch.bfloat16, group_size=64)
    table = torch.tensor(ref_lin.TABLE)
    wg = w.view(128, 8, 64)
    scale = (wg.abs().amax(-1) / table.abs().max()).clamp_min(1e-8).bfloat16()
    codes = ((wg / scale.float().unsqueeze(-1)).unsqueeze(-1) - table).abs().argmin(-1).view(128, 512)
    sz = torch.stack([scale.t(), torch.zeros_like(scale.t())], 2).contiguous()
    wd = dequant.dequant_lut(codes, table.bfloat16(), sz, 64, torch.bfloat16)
    assert _tol.frob_rel(y.cpu(), dequant.gemm(x, wd)) <= _tol.FROB_REL
    # and the quantised layer approximates the float one (4-bit: a few percent)
    full = (x.float() @ w.t()).bfloat16()
    assert _tol.frob_rel(y.cpu(), full) < 0.2


def test_mx4_linear_and_checkpoint(cuda_device):
    """MX4Linear against dequantize_mx4 + matmul, then a packed checkpoint round trip: the reloaded layer needs no
    re-packing and gives the same bits."""
    from any4_b200 import utils as host
    from any4_b200.modules import MX4Linear

    gen = torch.Generator().manual_seed(5)
    w = (torch.randn(64, 256, generator=gen) * 0.1).bfloat16()
    x = torch.randn(4, 256, generator=gen).bfloat16()
    lin = MX4Linear(256, 64, bias=True, device=cuda_device)
    lin.bias.data = torch.randn(64, generator=gen).bfloat16().to(cuda_device)
    lin.quantize_weight(w.to(cuda_device))
    y = lin(x.to(cuda_device))
    codes, exps = host.quantize_mx4(w, 32)
    wd = host.dequantize_mx4(codes, exps).to(torch.bfloat16)
    ref = (dequant.gemm(x, wd).float() + lin.bias.data.cpu().float()).bfloat16()
    assert y.shape == (4, 64) and _tol.frob_rel(y.cpu(), ref) <= 2e-3
    fresh = MX4Linear(256, 64, bias=True, device=cuda_device)
    fresh.load_state_dict(lin.state_dict())
    assert fresh.weight_reshaped and fresh.weight.dim() == 4
    assert torch.equal(fresh(x.to(cuda_device)), y)


def test_bind_host_zero_copy(cuda_device):
    """Any4Linear.bind_host / forward_host: pinned host activations in, pinned host outputs out, one kernel launch;
    the bits are those of the device-resident forward, and refilling the bound input buffer is all a new step needs."""
    from any4_b200.modules import Any4Linear

    gen = torch.Generator().manual_seed(9)
    lin = Any4Linear(1024, 256, bias=False, device=cuda_device, dtype=torch.bfloat16, group_size=128)
    _fill_any4(lin, gen)
    lin.reshape_weight(4)
    for m in (1, 3):
        xh = torch.randn(m, 1024, generator=gen).bfloat16().pin_memory()
        launch, yh = lin.bind_host(xh)
        for _ in range(3):
            xh.copy_(torch.randn(m, 1024, generator=gen).bfloat16())
            launch()
            torch.cuda.synchronize()
            assert torch.equal(yh.to(cuda_device), lin(xh.to(cuda_device)))
    y2 = lin.forward_host(xh)
    torch.cuda.synchronize()
    assert torch.equal(y2.to(cuda_device), lin(xh.to(cuda_device)))
    with pytest.raises(RuntimeError):
        lin.bind_host(torch.randn(1, 1024).bfloat16())           # not pinned
    with pytest.raises(RuntimeError):
        lin.bind_host(xh, out=torch.empty(1, 8).bfloat16().pin_memory())
