"""GPU parity against the UNMODIFIED reference kernels: tests/golden/golden_gpu.npz holds the
outputs of the reference extension (compiled for sm_100a by oracle/build_ref.py, run on a B200 by
`oracle/ref_runner.py golden`) on every case of oracle/cases.py.  Layout ops must be bit-exact;
GEMMs agree up to fp32 summation order before the final rounding (tests/_tol.py)."""
import os

import numpy as np
import pytest
import torch

from oracle import cases as C
from oracle import dequant
from tests import _tol

pytestmark = pytest.mark.gpu

GOLD_PATH = os.path.join(os.path.dirname(__file__), "golden", "golden_gpu.npz")
GOLD = np.load(GOLD_PATH) if os.path.exists(GOLD_PATH) else None
ALL = C.all_cases()


def _gold(case, key, dtype=None):
    name = f"{C.case_id(case)}/{key}"
    if GOLD is None or name not in GOLD:
        return None
    a = GOLD[name]
    if a.dtype == np.uint16:
        return torch.from_numpy(a.view(np.int16)).view(dtype)
    return torch.from_numpy(a)


@pytest.mark.skipif(GOLD is None, reason="golden_gpu.npz not generated yet")
@pytest.mark.parametrize("case", ALL, ids=[C.case_id(c) for c in ALL])
def test_matches_reference_kernel(case, cuda_device):
    import tinygemm  # noqa: F401

    inp = C.make_inputs(case)
    got = C.run_ops(case, inp, cuda_device)
    if case["kind"] == "convert":
        for key, g in got.items():
            ref = _gold(case, key, g.dtype)
            if ref is None:
                pytest.skip("case not in the golden file (rejected by the reference)")
            if case["op"] == "Aint8" and (-(-case["k"] // 16)) % case["ik"] != 0:
                pytest.skip("reference leaves the trailing inner k-tile uninitialised")
            assert g.shape == ref.shape
            if g.dtype in (torch.bfloat16, torch.float16):
                assert torch.equal(g.view(torch.int16), ref.view(torch.int16)), key
            else:
                assert torch.equal(g, ref), key
        return
    dt = C.DT[case["dt"]]
    ref = _gold(case, "y", dt)
    if ref is None:
        pytest.skip("case not in the golden file (rejected by the reference)")
    y = got["y"]
    assert y.shape == ref.shape
    # both kernels round the same fp32-accurate sum: they may differ by one ulp where the sums straddle a tie
    ulp = dequant.ulp_distance(y, ref)
    frac = (ulp == 0).double().mean().item()
    assert frac >= 0.95, f"only {frac:.3f} bit-equal to the reference kernel"
    assert _tol.frob_rel(y, ref) <= _tol.FROB_REL
    # where they differ, each must still be within the faithful bound of the exact result
    want = C.oracle_output(case, inp)
    assert _tol.frob_rel(ref, want["y"]) <= _tol.FROB_REL  # sanity: the golden itself matches the oracle
