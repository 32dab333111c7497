"""Host-side logic of the decode-step plumbing that needs no GPU: modules.fuse_rows tensor bookkeeping and the
argument checks of any4_b200.decode (the wrappers refuse CPU tensors: there is no CPU fallback)."""
import pytest
import torch


def _packed(n, k, g=128, seed=0):
    from any4_b200.modules import Any4Linear

    gen = torch.Generator().manual_seed(seed)
    lin = Any4Linear(k, n, bias=False, dtype=torch.bfloat16, group_size=g)
    lin.weight.data = torch.randint(-2**31, 2**31 - 1, (n // 8, k // 64, 32, 2), generator=gen, dtype=torch.int64).to(torch.int32)
    lin.lut.data = torch.randn(n, 16, generator=gen).bfloat16()
    lin.scales_and_zeros.data = torch.randn(k // g, n, 2, generator=gen).bfloat16()
    lin.weight_reshaped = True
    return lin


def test_fuse_rows_concatenates_packed_tensors():
    from any4_b200.modules import fuse_rows

    a, b, c = _packed(64, 256, seed=1), _packed(16, 256, seed=2), _packed(16, 256, seed=3)
    f = fuse_rows([a, b, c])
    assert (f.in_features, f.out_features, f.weight_reshaped, f.bias) == (256, 96, True, None)
    assert f.weight.shape == (12, 4, 32, 2) and f.weight.dtype == torch.int32
    assert torch.equal(f.weight[:8], a.weight) and torch.equal(f.weight[8:10], b.weight) and torch.equal(f.weight[10:], c.weight)
    assert f.scales_and_zeros.shape == (2, 96, 2) and f.scales_and_zeros.is_contiguous()
    assert torch.equal(f.scales_and_zeros[:, 64:80], b.scales_and_zeros)
    assert torch.equal(f.lut[80:], c.lut)


def test_fuse_rows_rejects_mismatches():
    from any4_b200.modules import Any4Linear, fuse_rows

    a = _packed(64, 256)
    with pytest.raises(ValueError):
        fuse_rows([a, _packed(64, 512)])                       # different in_features
    with pytest.raises(ValueError):
        fuse_rows([a, _packed(64, 256, g=64)])                 # different group size
    unpacked = Any4Linear(256, 64, bias=False, dtype=torch.bfloat16)
    with pytest.raises(ValueError):
        fuse_rows([a, unpacked])                               # not packed yet
    with pytest.raises(ValueError):
        fuse_rows([a, a], interleave=True)                     # interleaving needs UNPACKED layers
    glob = Any4Linear(256, 64, bias=False, dtype=torch.bfloat16, per_row=False)
    glob.weight_reshaped = True
    with pytest.raises(ValueError):
        fuse_rows([glob, glob])                                # one global LUT cannot be concatenated


def test_decode_wrappers_refuse_cpu_tensors():
    from any4_b200 import decode as D

    h = torch.zeros(1, 64, dtype=torch.bfloat16)
    with pytest.raises(RuntimeError):
        D.add_rmsnorm(h, None, torch.ones(64, dtype=torch.bfloat16), 1e-5)
    with pytest.raises(RuntimeError):
        D.silu_mul(torch.zeros(1, 128, dtype=torch.bfloat16))
