"""Parity at the headline sizes (BASELINE.json configs 1-4) against the UNMODIFIED reference kernels:
tests/golden/golden_big.npz holds 256 sampled output columns of the reference extension (compiled for sm_100a by
oracle/build_ref.py, run on a B200 by `oracle/ref_runner.py golden_big`) for any4 row-wise / nf4 / int4 / mx4 at
4096^2, 8192^2, 11008^2, 14336x4096 and 4096x14336, m in {1, 4, 8, 16}, both weight sides.  Covers what the small
cases cannot: persistent multi-row-block CTAs, stream-K splits with the cross-CTA fix-up, partially filled last
stages (k = 11008 = 10.75 stages) and the resident-activation limit (k = 14336 = 14 stages)."""
import os

import numpy as np
import pytest
import torch

from oracle import big_cases as B
from oracle import dequant
from tests import _tol

pytestmark = pytest.mark.gpu

GOLD_PATH = os.path.join(os.path.dirname(__file__), "golden", "golden_big.npz")
GOLD = np.load(GOLD_PATH) if os.path.exists(GOLD_PATH) else None
CASES = B.cases()


@pytest.mark.skipif(GOLD is None, reason="golden_big.npz not generated yet")
@pytest.mark.parametrize("case", CASES, ids=[B.case_id(c) for c in CASES])
def test_headline_shape_matches_reference_kernel(case, cuda_device):
    import tinygemm  # noqa: F401

    name = B.case_id(case)
    if name not in GOLD:
        pytest.skip("case not in the golden file (rejected by the reference)")
    ref = B.from_u16(GOLD[name])
    y = B.run_ops(case, cuda_device)
    assert y.shape == ref.shape and y.dtype == torch.bfloat16
    assert torch.isfinite(y.float()).all()
    # both kernels round the same fp32-accurate sum: they may differ by one ulp where the sums straddle a tie
    ulp = dequant.ulp_distance(y, ref)
    frac = (ulp == 0).double().mean().item()
    assert frac >= 0.95, f"only {frac:.3f} bit-equal to the reference kernel"
    # (no per-element ulp bound: where a sum cancels to ~0 a difference in the fp32 summation order is many ulps of the
    # tiny result; the Frobenius bound below is relative to the whole sample)
    assert _tol.frob_rel(y, ref) <= _tol.FROB_REL  # the north-star tolerance: 1e-3 relative


GOLD_W8_PATH = os.path.join(os.path.dirname(__file__), "golden", "golden_big_w8.npz")
GOLD_W8 = np.load(GOLD_W8_PATH) if os.path.exists(GOLD_W8_PATH) else None
CASES_W8 = B.cases_w8()


@pytest.mark.skipif(GOLD_W8 is None, reason="golden_big_w8.npz not generated yet")
@pytest.mark.parametrize("case", CASES_W8, ids=[B.case_id(c) for c in CASES_W8])
def test_headline_shape_w8_matches_reference_kernel(case, cuda_device):
    """int8 (ring kernel: both layouts, one pass and two passes of activation rows, 4 / 8 / 10.75 chunks per row tile)
    and 16-bit weights at the headline sizes against sampled outputs of the reference's kernels
    (`oracle/ref_runner.py golden_big_w8`)."""
    import tinygemm  # noqa: F401

    name = B.case_id(case)
    if name not in GOLD_W8:
        pytest.skip("case not in the golden file (rejected by the reference)")
    ref = B.from_u16(GOLD_W8[name])
    y = B.run_ops_w8(case, cuda_device)
    assert y.shape == ref.shape and y.dtype == torch.bfloat16
    assert torch.isfinite(y.float()).all()
    frac = (dequant.ulp_distance(y, ref) == 0).double().mean().item()
    assert frac >= 0.95, f"only {frac:.3f} bit-equal to the reference kernel"
    assert _tol.frob_rel(y, ref) <= _tol.FROB_REL


@pytest.mark.parametrize("fmt", ["any4r", "int4", "mx4"])
def test_headline_shape_matches_oracle(fmt, cuda_device):
    """The sampled columns of the 4096^2 m = 4 case against the float64 CPU restatement (independent of the golden)."""
    import tinygemm  # noqa: F401
    from oracle import layouts

    n = k = 4096
    case = dict(fmt=fmt, side="right", n=n, k=k, m=4)
    y = B.run_ops(case, cuda_device)
    s = B.shape_inputs(n, k)
    cols = B.sample_cols(n)
    packed = s["words"].view(n // 8, k // 64, 32, 2)
    tiles = (cols // 8).unique()
    codes = torch.from_numpy(layouts.from_Bint4(packed[tiles].numpy()))  # [len(tiles) * 8][k]
    pos = {int(t): i for i, t in enumerate(tiles)}
    rows = torch.tensor([pos[int(c) // 8] * 8 + int(c) % 8 for c in cols])
    codes = codes[rows].to(torch.int64)
    if fmt == "mx4":
        w = dequant.dequant_mx4(codes, s["exps"][cols], B.MX_G)
    elif fmt == "int4":
        w = dequant.dequant_int4(codes, s["sz"][:, cols].contiguous(), B.G, torch.bfloat16)
    else:
        w = dequant.dequant_lut(codes, s["lut"][cols], s["sz"][:, cols].contiguous(), B.G, torch.bfloat16)
    x = s["x"][:4]
    y64 = dequant.gemm_f64(x, w)
    absdot = x.double().abs() @ w.double().abs().t()
    nbad, worst = _tol.check_faithful(y, y64, absdot, torch.bfloat16)
    assert nbad == 0, f"{nbad} elements outside the faithful-rounding bound (worst {worst:.2f}x)"
