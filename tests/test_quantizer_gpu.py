"""GPU: the any4 quantizer kernel (tg_quantize_any4_rows through any4_b200.quantize) against the CPU oracle
(oracle/quantizer.py, pinned to the reference's group_q / run_kmeans) and against the committed reference goldens."""
import os

import numpy as np
import pytest
import torch

from oracle import quantizer as Q

pytestmark = pytest.mark.gpu

GOLD = np.load(os.path.join(os.path.dirname(__file__), "golden", "golden_quantizer.npz"))


def bf16(name):
    return torch.from_numpy(GOLD[name + "__bf16"].view(np.int16).copy()).view(torch.bfloat16)


@pytest.mark.parametrize("c", range(int(GOLD["n_cases"])))
def test_kernel_vs_reference_golden(c, cuda_device):
    from any4_b200.quantize import anyq_quantize_tensor

    n, k, g, weighted = (int(v) for v in GOLD[f"q{c}_meta"])
    W = bf16(f"q{c}_w")
    sw = torch.from_numpy(GOLD[f"q{c}_sw"]) if weighted else None
    assign, any4, sz, lut = anyq_quantize_tensor(W.to(cuda_device), q_group_size=g, sample_weight=sw, return_lut=True)
    assert torch.equal(sz.cpu(), bf16(f"q{c}_sz"))                       # group statistics: bit-exact
    labels = torch.from_numpy(GOLD[f"q{c}_labels"])
    assert (assign.cpu() == labels).float().mean() >= 0.998               # the reference's own Lloyd, same init
    cen = torch.from_numpy(GOLD[f"q{c}_centroids"])
    assert (any4.cpu().float() - cen).abs().max() <= 2.0 ** -7 * 15        # centroids (stored in bf16)
    assert torch.equal(lut, any4 - 8)                                     # quantize.py:893, in the weight dtype
    Wd = Q.dequantize(assign.cpu(), any4.cpu(), sz.cpu(), g)
    mse = float(((Wd - W.float()) ** 2).mean())
    if not weighted:
        assert mse <= 1.02 * float(GOLD[f"q{c}_sklearn_mse"])              # quality of the default (sklearn) path


@pytest.mark.parametrize("n,k,g,dt", [(64, 4096, 128, torch.bfloat16), (24, 1024, 32, torch.float16), (8, 11008, 128, torch.bfloat16)])
def test_kernel_vs_oracle_and_packing(n, k, g, dt, cuda_device):
    import tinygemm  # noqa: F401  (registers torch.ops.tinygemm)
    from any4_b200.quantize import anyq_quantize_tensor

    gen = torch.Generator().manual_seed(n + k)
    W = (torch.randn(n, k, generator=gen) * 0.03).to(dt)
    want = Q.quantize_any4(W, g)
    assign, any4, sz, packed = anyq_quantize_tensor(W.to(cuda_device), q_group_size=g, pack_inner_k=4)
    assert torch.equal(sz.cpu(), want["sz"])
    assert (assign.cpu() == want["codes"]).float().mean() >= 0.998
    assert (any4.cpu().float() - want["any4"].float()).abs().max() <= (2.0 ** -7 if dt == torch.bfloat16 else 2.0 ** -9) * 15
    # the packed output IS the convert op applied to the returned codes (bit-exact), for every inner-k
    ops = torch.ops.tinygemm
    assert torch.equal(packed, ops.convert_matrix_to_m16n8k16_Bint4_layout(assign, 4))
    for ik in (2, 8):
        a2, _, _, p2 = anyq_quantize_tensor(W.to(cuda_device), q_group_size=g, pack_inner_k=ik)
        assert torch.equal(a2, assign)                                      # deterministic
        assert torch.equal(p2, ops.convert_matrix_to_m16n8k16_Bint4_layout(a2, ik))


def test_any4_linear_from_float(cuda_device):
    from any4_b200.quantize import any4_linear_from_float

    gen = torch.Generator().manual_seed(7)
    lin = torch.nn.Linear(1024, 256, bias=True, dtype=torch.bfloat16)
    with torch.no_grad():
        lin.weight.copy_((torch.randn(256, 1024, generator=gen) * 0.05).bfloat16())
    lin = lin.to(cuda_device)
    q = any4_linear_from_float(lin, group_size=128)
    assert q.weight_reshaped and q.weight.shape == (32, 16, 32, 2) and q.lut.shape == (256, 16)
    x = torch.randn(5, 1024, generator=gen).bfloat16().to(cuda_device)
    with torch.no_grad():
        y, want = q(x).float(), lin(x).float()
    rel = float((y - want).norm() / want.norm())
    assert rel < 0.12, rel            # 16 optimal levels on Gaussian weights: ~0.1 relative (Lloyd-Max: MSE = 0.0095 sigma^2)
    # and better than plain int4 with the same groups
    from any4_b200 import utils as U
    codes, sz = U.group_quantize_tensor(lin.weight.data, 4, 128)
    wi = ((codes.float() - 8) * sz[..., 0].t().float().repeat_interleave(128, 1) + sz[..., 1].t().float().repeat_interleave(128, 1))
    with torch.no_grad():
        rel_int4 = float((x.float() @ wi.t() + lin.bias.float() - want).norm() / want.norm())
    assert rel < rel_int4


def test_rejects_what_it_does_not_cover(cuda_device):
    from any4_b200.quantize import anyq_quantize_tensor

    W = torch.randn(8, 256, device=cuda_device).bfloat16()
    with pytest.raises(NotImplementedError):
        anyq_quantize_tensor(W, n_bit=3)
    with pytest.raises(RuntimeError):
        anyq_quantize_tensor(W.float())
    with pytest.raises(RuntimeError):
        anyq_quantize_tensor(W, q_group_size=48)
