"""Decode-step plumbing (any4_b200.decode, modules.fuse_rows) against plain torch references on the GPU."""
import pytest
import torch
import torch.nn.functional as TF

pytestmark = pytest.mark.gpu

HEADS, KV, HD = 32, 8, 128


def _dev():
    return torch.device("cuda:0")


@pytest.mark.parametrize("dtype", [torch.bfloat16, torch.float16])
@pytest.mark.parametrize("n", [4096, 8192, 1024])
@pytest.mark.parametrize("with_delta", [True, False])
def test_add_rmsnorm(dtype, n, with_delta):
    from any4_b200 import decode as D

    g = torch.Generator(device=_dev()).manual_seed(n)
    h = torch.randn(1, n, device=_dev(), generator=g).to(dtype)
    delta = torch.randn(1, n, device=_dev(), generator=g).to(dtype) if with_delta else None
    w = (torch.rand(n, device=_dev(), generator=g) + 0.5).to(dtype)
    h_ref = (h + delta) if with_delta else h.clone()      # rounded residual add, as the framework does
    ref = TF.rms_norm(h_ref.float(), (n,), w.float(), 1e-5)
    h_run = h.clone()
    out = D.add_rmsnorm(h_run, delta, w, 1e-5)
    assert torch.equal(h_run, h_ref)                      # the residual stream is updated in place, bit-exact
    ulp = 2.0 ** (-8 if dtype == torch.bfloat16 else -11)
    assert ((out.float() - ref).abs() <= ulp * ref.abs() + 1e-6).all()   # one rounding of the fp32 result


@pytest.mark.parametrize("dtype", [torch.bfloat16, torch.float16])
def test_silu_mul(dtype):
    from any4_b200 import decode as D

    n = 14336
    g = torch.Generator(device=_dev()).manual_seed(3)
    gu = (torch.randn(1, 2 * n, device=_dev(), generator=g) * 2).to(dtype)
    out = D.silu_mul(gu)
    ref = TF.silu(gu[:, :n]) * gu[:, n:]                  # two rounded framework ops
    ulp = 2.0 ** (-7 if dtype == torch.bfloat16 else -10)
    assert out.shape == (1, n)
    assert ((out.float() - ref.float()).abs() <= ulp * ref.float().abs() + 1e-7).all()
    assert (out == ref).float().mean() > 0.99             # same roundings: almost always identical


@pytest.mark.parametrize("dtype", [torch.bfloat16, torch.float16])
@pytest.mark.parametrize("ctx", [0, 1, 128, 511])
def test_rope_attention(dtype, ctx):
    from any4_b200 import decode as D

    dev = _dev()
    g = torch.Generator(device=dev).manual_seed(ctx + 1)
    qkv = torch.randn(1, (HEADS + 2 * KV) * HD, device=dev, generator=g).to(dtype)
    kc = torch.randn(1, KV, ctx + 1, HD, device=dev, generator=g).to(dtype)
    vc = torch.randn(1, KV, ctx + 1, HD, device=dev, generator=g).to(dtype)
    inv = 1.0 / (500000.0 ** (torch.arange(0, HD, 2, device=dev).float() / HD))
    ang = torch.cat([ctx * inv, ctx * inv])
    cos, sin = ang.cos().to(dtype), ang.sin().to(dtype)

    # reference: the stock ops of bench_llama.Block
    q = qkv[:, : HEADS * HD].view(1, HEADS, 1, HD)
    k = qkv[:, HEADS * HD: (HEADS + KV) * HD].view(1, KV, 1, HD)
    v = qkv[:, (HEADS + KV) * HD:].view(1, KV, 1, HD)

    def rope(t):
        t1, t2 = t[..., : HD // 2], t[..., HD // 2:]
        return t * cos.view(1, 1, 1, HD) + torch.cat((-t2, t1), -1) * sin.view(1, 1, 1, HD)

    kc_ref, vc_ref = kc.clone(), vc.clone()
    kc_ref[:, :, ctx:] = rope(k)
    vc_ref[:, :, ctx:] = v
    ref = TF.scaled_dot_product_attention(rope(q).float(), kc_ref.float(), vc_ref.float(), enable_gqa=True).reshape(1, HEADS * HD)

    kc_run, vc_run = kc.clone(), vc.clone()
    out = D.rope_attention(qkv, cos, sin, kc_run, vc_run, ctx, HEADS, KV, HD)
    assert torch.equal(kc_run, kc_ref) and torch.equal(vc_run, vc_ref)   # cache append incl. the rotated k, bit-exact
    err = (out.float() - ref).abs().max().item()
    assert err <= (2e-2 if dtype == torch.bfloat16 else 4e-3) * max(1.0, ref.abs().max().item()), err


def test_argument_checks():
    from any4_b200 import decode as D

    dev = _dev()
    h = torch.zeros(1, 4100, device=dev, dtype=torch.bfloat16)
    with pytest.raises(RuntimeError):
        D.add_rmsnorm(h, None, torch.ones(4100, device=dev, dtype=torch.bfloat16), 1e-5)   # n % 8
    qkv = torch.zeros(1, (HEADS + 2 * KV) * HD, device=dev, dtype=torch.bfloat16)
    c = torch.zeros(HD, device=dev, dtype=torch.bfloat16)
    kc = torch.zeros(1, KV, 4, HD, device=dev, dtype=torch.bfloat16)
    with pytest.raises(RuntimeError):
        D.rope_attention(qkv, c, c, kc, kc.clone(), 4, HEADS, KV, HD)                      # pos >= cache_len


@pytest.mark.parametrize("rows", [(4096, 1024, 1024), (512, 256), (6144, 4096)])
def test_fuse_rows(rows):
    """One launch over concatenated rows computes every row as the separate layers do: same dequantised weights, same
    exact products; the kernel choice and the stream-K split depend on the launch size, so the ORDER of the fp32
    partial sums (and with it the last bf16 bit of a few outputs) may differ."""
    from any4_b200.modules import Any4Linear, fuse_rows
    from bench import G, synth_layer

    dev = _dev()
    k = 4096
    lins = []
    for i, n in enumerate(rows):
        lin = Any4Linear(k, n, bias=False, device=dev, dtype=torch.bfloat16, group_size=G)
        w, lut, sz = synth_layer(n, k, 40 + i, dev)
        lin.weight.data, lin.lut.data, lin.scales_and_zeros.data = w, lut, sz
        lin.weight_reshaped = True
        lins.append(lin)
    fused = fuse_rows(lins)
    for m in (1, 3):
        x = torch.randn(m, k, device=dev, generator=torch.Generator(device=dev).manual_seed(m)).bfloat16()
        got, want = fused(x), torch.cat([lin(x) for lin in lins], -1)
        # (an output that cancels to ~0 may differ by many of ITS ulps: bound relative to the row's scale as well)
        assert ((got.float() - want.float()).abs() <= 2.0 ** -7 * want.float().abs() + 2.0 ** -9 * want.float().abs().max()).all()
        assert (got == want).float().mean() > 0.98


@pytest.mark.parametrize("m", [1, 3])
@pytest.mark.parametrize("n,k", [(4096, 4096), (512, 4096), (2048, 1024)])
def test_linear_silu_pairs(m, n, k):
    """gate|up row-interleaved + activation in the GEMV epilogue == separate GEMVs followed by silu * mul."""
    import copy

    from any4_b200 import decode as D
    from any4_b200.modules import Any4Linear, fuse_rows

    dev = _dev()
    g = torch.Generator(device=dev).manual_seed(n + k + m)
    lins = []
    for _ in range(2):
        lin = Any4Linear(k, n, bias=False, device=dev, dtype=torch.bfloat16, group_size=128)
        lin.weight.data = torch.randint(0, 16, (n, k), device=dev, generator=g, dtype=torch.int32)
        lin.lut.data = ((torch.rand(n, 16, device=dev, generator=g) * 15).sort(1).values.bfloat16() - 8)
        lin.scales_and_zeros.data = torch.stack([torch.rand(k // 128, n, device=dev, generator=g) * 0.02 + 0.001,
                                                 torch.randn(k // 128, n, device=dev, generator=g) * 0.02], 2).bfloat16()
        lins.append(lin)
    fused = fuse_rows([copy.deepcopy(l) for l in lins], interleave=True)
    for lin in lins:
        lin.reshape_weight(4)
    x = torch.randn(m, k, device=dev, generator=g).bfloat16()
    yg, yu = lins[0](x), lins[1](x)
    want = TF.silu(yg) * yu
    plain = fused(x).view(m, n, 2)                                 # the interleaved weight through the plain GEMV
    for a, b in ((plain[..., 0], yg), (plain[..., 1], yu)):        # same weights, possibly another summation order
        assert ((a.float() - b.float()).abs() <= 2.0 ** -7 * b.float().abs() + 2.0 ** -9 * b.float().abs().max()).all()
        assert (a == b).float().mean() > 0.98
    got = D.linear_silu_pairs(fused, x)
    assert got.shape == (m, n)
    assert ((got.float() - want.float()).abs() <= 2.0 ** -6 * want.float().abs() + 2.0 ** -8 * want.float().abs().max()).all()
    assert (got == want).float().mean() > 0.97
    # the fused epilogue IS the plain GEMV + tg_decode_silu_mul, bit for bit
    for r in range(m):
        assert torch.equal(got[r:r + 1], D.silu_mul(torch.cat([plain[r:r + 1, :, 0], plain[r:r + 1, :, 1]], -1).contiguous()))
