"""GPU parity: the eight layout ops, bit-exact against the numpy restatement (oracle/layouts.py)
and round-trips (reference: tests/tinygemm/test_tinygemm_convert.py:20-96)."""
import pytest
import torch

from oracle import cases as C

pytestmark = pytest.mark.gpu

CASES = C.convert_cases()


@pytest.mark.parametrize("case", CASES, ids=[C.case_id(c) for c in CASES])
def test_convert_bit_exact(case, cuda_device):
    import tinygemm  # noqa: F401

    inp = C.make_inputs(case)
    got = C.run_ops(case, inp, cuda_device)
    want = C.oracle_output(case, inp)
    for key, ref in want.items():
        g = got[key]
        assert g.shape == ref.shape, (key, g.shape, ref.shape)
        if g.dtype in (torch.bfloat16, torch.float16):
            assert torch.equal(g.view(torch.int16), ref.view(torch.int16)), key
        else:
            assert torch.equal(g, ref), key


def test_convert_errors(cuda_device):
    import tinygemm  # noqa: F401

    ops = torch.ops.tinygemm
    codes = torch.zeros(8, 96, dtype=torch.int32, device=cuda_device)
    with pytest.raises(RuntimeError):
        ops.convert_matrix_to_m16n8k16_Bint4_layout(codes, 4)  # 96 % 64 != 0
    with pytest.raises(RuntimeError):
        ops.convert_matrix_to_m16n8k16_Bint4_layout(codes, 3)
    with pytest.raises(RuntimeError):
        ops.convert_matrix_to_m16n8k16_Aint4_layout(codes.float(), 1)
    with pytest.raises(RuntimeError):
        ops.convert_matrix_to_m16n8k16_A_layout(torch.zeros(4, 4, device=cuda_device), 1)  # fp32
    with pytest.raises(RuntimeError):
        ops.convert_matrix_to_m16n8k16_A_layout(torch.zeros(4, 4, dtype=torch.bfloat16), 1)  # CPU tensor


def test_dequant_int4_debug_op(cuda_device):
    import tinygemm  # noqa: F401

    w = torch.randint(-2**31, 2**31 - 1, (1000,), dtype=torch.int64).to(torch.int32)
    out = torch.ops.tinygemm.tinygemm_dequant_int4(w.to(cuda_device)).cpu().float().view(-1, 8)
    u = w.to(torch.int64) & 0xFFFFFFFF
    shifts = [0, 16, 4, 20, 8, 24, 12, 28]
    want = torch.stack([((u >> s) & 0xF).float() - 8 for s in shifts], dim=1)
    assert torch.equal(out, want)


@pytest.mark.parametrize("ik_a", [1, 2, 4])
@pytest.mark.parametrize("ik_b", [2, 4, 8])
def test_repack_Aint4_to_Bint4(ik_a, ik_b, cuda_device):
    """tg_repack_Aint4_to_Bint4: the packed A layout of a code matrix -> its packed B layout, bit for bit (both layouts
    hold the same nibbles; the B layout of the reference's own convert op is the oracle)."""
    import ctypes

    import tinygemm  # noqa: F401
    from any4_b200 import _native

    ops, lib = torch.ops.tinygemm, _native.capi()
    gen = torch.Generator().manual_seed(ik_a * 10 + ik_b)
    for n, k in ((48, 256), (16, 1024), (80, 512)):
        codes = torch.randint(0, 16, (n, k), generator=gen, dtype=torch.int32).to(cuda_device)
        a = ops.convert_matrix_to_m16n8k16_Aint4_layout(codes, ik_a)
        want = ops.convert_matrix_to_m16n8k16_Bint4_layout(codes, ik_b)
        rows = a.shape[0] * 16
        got = torch.empty((rows // 8, k // (16 * ik_b), 32, ik_b // 2), dtype=torch.int32, device=cuda_device)
        rc = lib.tg_repack_Aint4_to_Bint4(ctypes.c_void_p(a.data_ptr()), ctypes.c_void_p(got.data_ptr()), rows, k, ik_a, ik_b,
                                          ctypes.c_void_p(torch.cuda.current_stream().cuda_stream))
        assert rc == 0, _native.last_error()
        assert torch.equal(got[: want.shape[0]], want)
        assert int(got[want.shape[0]:].abs().sum()) == 0   # rows padded to 16 are zero codes
