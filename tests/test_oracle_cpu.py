"""CPU: pins the oracle and the host-side quantizers against golden vectors produced by importing
the reference itself (tests/golden/make_golden_cpu.py -> golden_cpu.npz), and checks the layout
restatement's internal consistency."""
import os

import numpy as np
import pytest
import torch

from any4_b200 import utils as host
from oracle import cpu_path, dequant, layouts

GOLD = np.load(os.path.join(os.path.dirname(__file__), "golden", "golden_cpu.npz"))


def g(name):
    if name + "__bf16" in GOLD:
        a = GOLD[name + "__bf16"]
        return torch.from_numpy(a.view(np.int16)).view(torch.bfloat16)
    return torch.from_numpy(GOLD[name])


def test_group_quantize_matches_reference():
    for i in range(int(GOLD["gq_cases"])):
        n_bit, grp = (int(v) for v in GOLD[f"gq{i}_meta"])
        w = g(f"gq{i}_w")
        codes, sz = host.group_quantize_tensor(w, n_bit, grp)
        assert torch.equal(codes, g(f"gq{i}_codes"))
        ref = g(f"gq{i}_sz")
        assert sz.dtype == ref.dtype and torch.equal(sz.float(), ref.float())
    codes, sz = host.group_quantize_tensor(torch.eye(128, dtype=torch.bfloat16), 4, 32)
    assert torch.equal(codes, g("gq_eye_codes")) and torch.equal(sz.float(), g("gq_eye_sz").float())


def test_mx4_matches_reference():
    for i in range(int(GOLD["mx_cases"])):
        x = g(f"mx{i}_x")
        q, e = host.quantize_mx4(x, 32)
        assert torch.equal(q, g(f"mx{i}_q"))
        assert torch.equal(e, g(f"mx{i}_e"))
        assert torch.equal(host.dequantize_mx4(q, e), g(f"mx{i}_d"))
        # the GPU numerics restatement agrees with the reference's fp32 dequantize when cast to bf16
        d = dequant.dequant_mx4(q, e, 32, torch.bfloat16)
        assert torch.equal(d.float(), g(f"mx{i}_d").to(torch.bfloat16).float())


def test_cpu_path_matches_reference():
    for i in range(int(GOLD["cpu_cases"])):
        per_row, grp = (int(v) for v in GOLD[f"cpu{i}_meta"])
        w = cpu_path.anyq_dequantize(g(f"cpu{i}_assign"), g(f"cpu{i}_any4"), g(f"cpu{i}_sz"), 4, grp, bool(per_row))
        assert torch.equal(w.float(), g(f"cpu{i}_w").float())
        y = cpu_path.any4_linear_forward(g(f"cpu{i}_x"), g(f"cpu{i}_assign"), g(f"cpu{i}_any4"), g(f"cpu{i}_sz"),
                                         grp, bool(per_row))
        assert torch.equal(y.float(), g(f"cpu{i}_y").float())


def test_fma_rn_single_rounding():
    # cases where separate rounding of the product differs from the fused result
    v = torch.tensor([3.0, -7.0, 5.0, 1.0], dtype=torch.bfloat16)
    s = torch.tensor([0.0133, 0.00787, 0.0101, 1.0], dtype=torch.bfloat16)
    z = torch.tensor([0.00411, -0.0021, 1e-5, -1.0], dtype=torch.bfloat16)
    got = dequant.fma_rn(v, s, z, torch.bfloat16)
    exact = v.double() * s.double() + z.double()  # exact in float64 for these magnitudes
    want = dequant._f64_to_bf16_torch(exact.numpy())
    assert torch.equal(got.float(), want.float())
    # fp16 path
    got16 = dequant.fma_rn(v.half(), s.half(), z.half(), torch.float16)
    exact16 = v.half().double() * s.half().double() + z.half().double()
    assert torch.equal(got16.float(), torch.from_numpy(exact16.numpy().astype(np.float16)).float())


def test_int4_identity_fixture_is_exact():
    """W = I: code 0 -> -8*s + 8*s = 0 and code 15 -> 15*bf16(1/15)+... = 1 exactly in bf16"""
    codes, sz = host.group_quantize_tensor(torch.eye(64, dtype=torch.bfloat16), 4, 32)
    w = dequant.dequant_int4(codes, sz, 32, torch.bfloat16)
    assert torch.equal(w.float(), torch.eye(64))


@pytest.mark.parametrize("ik", [1, 2, 4])
def test_layout_roundtrip_Aint4(ik):
    rng = np.random.default_rng(ik)
    codes = rng.integers(0, 16, size=(37, 200), dtype=np.int32)
    packed = layouts.to_Aint4(codes, ik)
    assert packed.shape == (3, -(-200 // (16 * ik)), 32, ik)
    back = layouts.from_Aint4(packed)
    assert np.array_equal(back[:37, :200], codes) and back[37:].sum() == 0 and back[:, 200:].sum() == 0


@pytest.mark.parametrize("ik", [2, 4, 8])
def test_layout_roundtrip_Bint4(ik):
    rng = np.random.default_rng(ik)
    codes = rng.integers(0, 16, size=(21, 256), dtype=np.int32)
    packed = layouts.to_Bint4(codes, ik)
    assert packed.shape == (3, 256 // (16 * ik), 32, ik // 2)
    assert np.array_equal(layouts.from_Bint4(packed)[:21], codes)


def test_layout_roundtrip_int8_and_16bit():
    rng = np.random.default_rng(0)
    codes = rng.integers(0, 256, size=(24, 128), dtype=np.int32)
    for ik in (1, 2):
        assert np.array_equal(layouts.from_Aint8(layouts.to_Aint8(codes, ik))[:24, :128], codes)
    for ik in (1, 2, 4):
        assert np.array_equal(layouts.from_Bint8(layouts.to_Bint8(codes, ik))[:24], codes)
    x = rng.integers(0, 65536, size=(19, 45)).astype(np.uint16)
    assert np.array_equal(layouts.from_A(layouts.to_A(x), 19, 45), x)
    for ik in (1, 2):
        assert np.array_equal(layouts.from_B(layouts.to_B(x, ik), 19, 45), x)


def test_packed_word_nibble_order():
    # one B tile pair: word of lane t = tile0 (v0..3) | tile1 (v4..7), order v7 v5 v3 v1 v6 v4 v2 v0
    codes = np.arange(8 * 32, dtype=np.int32).reshape(8, 32) % 16
    w = layouts.to_Bint4(codes, 2).view(np.uint32)[0, 0, :, 0]
    t = 5  # g = 1, q = 1 -> row 1, k0 = 2
    vals = [codes[1, 2], codes[1, 3], codes[1, 10], codes[1, 11], codes[1, 18], codes[1, 19], codes[1, 26], codes[1, 27]]
    want = 0
    for v, s in zip(vals, [0, 16, 4, 20, 8, 24, 12, 28]):
        want |= int(v) << s
    assert int(w[t]) == want
