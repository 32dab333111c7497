"""CPU: module construction, the checkpoint format and the fixed-table module shells - no kernels run."""
import torch

from any4_b200.modules import Any4Linear, FP4Linear, Int4Linear, Int8Linear, MX4Linear, NF4Linear


def _fake_pack(lin, shape, inner_k):
    lin.weight = torch.nn.Parameter(torch.arange(int(torch.tensor(shape).prod()), dtype=torch.int32).view(*shape), requires_grad=False)
    lin.weight_reshaped, lin.w_inner_k = True, inner_k


def test_state_dict_keys_match_reference_plus_extra_state():
    """modules.py:12-230 parameter names; the only addition is the extra-state entry."""
    assert list(Int4Linear(256, 64, dtype=torch.bfloat16).state_dict()) == ["weight", "scales_and_zeros", "bias", "_extra_state"]
    assert list(Int8Linear(256, 64, dtype=torch.bfloat16).state_dict()) == ["weight", "scales_and_zeros", "bias", "_extra_state"]
    assert list(Any4Linear(256, 64, dtype=torch.bfloat16).state_dict()) == ["weight", "scales_and_zeros", "lut", "bias", "_extra_state"]
    assert list(MX4Linear(256, 64, bias=False).state_dict()) == ["weight", "exponents", "_extra_state"]


def test_packed_checkpoint_round_trip():
    """A packed layer saves its packed weight + how it was packed; a freshly constructed layer loads it as is
    (the reference loses `weight_reshaped` / `w_inner_k`, modules.py:194)."""
    src = Any4Linear(256, 64, bias=True, dtype=torch.bfloat16, per_row=True)
    torch.nn.init.normal_(src.lut), torch.nn.init.normal_(src.scales_and_zeros), torch.nn.init.normal_(src.bias)
    _fake_pack(src, (8, 2, 32, 4), 8)
    buf = __import__("io").BytesIO()
    torch.save(src.state_dict(), buf)
    buf.seek(0)
    dst = Any4Linear(256, 64, bias=True, dtype=torch.bfloat16, per_row=True)
    assert not dst.weight_reshaped and dst.weight.shape == (64, 256)
    res = dst.load_state_dict(torch.load(buf))
    assert not res.missing_keys and not res.unexpected_keys
    assert dst.weight_reshaped and dst.w_inner_k == 8 and dst.kernel == src.kernel
    for k, v in src.state_dict().items():
        if k != "_extra_state":
            assert torch.equal(v, dst.state_dict()[k])


def test_reference_checkpoints_still_load():
    """No extra state in the file (a reference state_dict): strict loading works; an unpacked weight leaves the flags
    alone, a 4-D weight is recognised as packed and its inner-k recovered from the shape."""
    plain = {k: v for k, v in Int4Linear(256, 64, dtype=torch.bfloat16).state_dict().items() if k != "_extra_state"}
    dst = Int4Linear(256, 64, dtype=torch.bfloat16)
    dst.load_state_dict(plain)
    assert not dst.weight_reshaped
    packed = dict(plain, weight=torch.zeros(4, 4, 32, 4, dtype=torch.int32))  # A int4 layout, inner_k 4 (the class default kernel)
    dst = Int4Linear(256, 64, dtype=torch.bfloat16)
    dst.load_state_dict(packed)
    assert dst.weight_reshaped and dst.w_inner_k == 4
    packed8 = {k: v for k, v in Int8Linear(256, 64, dtype=torch.bfloat16).state_dict().items() if k != "_extra_state"}
    packed8["weight"] = torch.zeros(4, 8, 32, 4, dtype=torch.int32)               # A int8 layout, inner_k 2
    dst8 = Int8Linear(256, 64, dtype=torch.bfloat16)
    dst8.load_state_dict(packed8)
    assert dst8.weight_reshaped and dst8.w_inner_k == 2


def test_global_lut_checkpoint_into_per_row_module():
    src = Any4Linear(256, 64, bias=False, dtype=torch.bfloat16, per_row=False)
    dst = Any4Linear(256, 64, bias=False, dtype=torch.bfloat16, per_row=True)
    dst.load_state_dict(src.state_dict())
    assert dst.lut.shape == (16,)


def test_fixed_table_modules():
    nf4 = NF4Linear(256, 64, bias=False, dtype=torch.bfloat16)
    assert nf4.lut.shape == (16,) and not nf4.per_row and nf4.kernel == "linear_y_f16RM_x_f16RM_W_any4TC"
    assert float(nf4.lut[0]) == -1.0 and float(nf4.lut[15]) == 1.0 and float(nf4.lut[7]) == 0.0
    fp4 = FP4Linear(256, 64, bias=False, dtype=torch.float16)
    assert sorted(fp4.lut.abs().unique().tolist()) == [0.0, 0.5, 1.0, 1.5, 2.0, 3.0, 4.0, 6.0]
    mx = MX4Linear(256, 64, bias=True)
    assert mx.exponents.shape == (64, 8) and mx.exponents.dtype == torch.uint8
    try:
        MX4Linear(256, 64, dtype=torch.float16)
        raise AssertionError("fp16 must be rejected")
    except ValueError:
        pass
