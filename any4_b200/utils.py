"""Host-side data generators for the tinygemm path (drop-in for `tinygemm_lib.utils`).

Same names, argument meaning and return layouts as the reference
(tinygemm_lib/utils.py:27-67 group_quantize_tensor, :69-82 expand/extract helpers,
:85-134 round_to_mx4, :137-191 quantize_mx4, :194-232 dequantize_mx4; the MX rounding
rules restate tinygemm_lib/mx/mx_ops.py:52-125 and mx/elemwise_ops.py:85-200 for the
single format tinygemm uses, fp4 e2m1 with an e8m0 shared exponent).  These are
torch-level quantizers that run on whatever device their input lives on; they produce
the tensors the CUDA GEMV consumes and are not themselves on the decode path.
"""
import torch

# fp4 e2m1 magnitudes by code (sign bit = code & 8); FloatDefs.cuh:18-34 in the reference
MX4_VALUES = (0.0, 0.5, 1.0, 1.5, 2.0, 3.0, 4.0, 6.0, -0.0, -0.5, -1.0, -1.5, -2.0, -3.0, -4.0, -6.0)

_FP4_EMAX = 2          # largest normal exponent of e2m1
_FP4_MAX = 6.0         # largest normal magnitude
_E8M0_EMAX = 127


def group_quantize_tensor(w_orig, n_bit, q_group_size=128):
    """Asymmetric uniform group quantization along k.

    Returns (codes int32 [n][k] in [0, 2^n_bit-1], scales_and_zeros [k/g][n][2] in
    w_orig.dtype) with reconstruction  (code - 2^(n_bit-1)) * scale + zero.
    """
    assert q_group_size > 1
    assert w_orig.dim() == 2
    assert w_orig.shape[-1] % q_group_size == 0
    w = w_orig.float()
    n, k = w.shape
    grouped = w.reshape(-1, q_group_size)
    assert not torch.isnan(grouped).any()

    hi = grouped.amax(dim=1, keepdim=True)
    lo = grouped.amin(dim=1, keepdim=True)
    levels = 2**n_bit - 1
    scales = (hi - lo).clamp(min=1e-6) / levels
    zeros = lo + scales * (2 ** (n_bit - 1))

    codes = grouped.sub(lo).div(scales).round().clamp_(0, levels)
    codes = codes.to(torch.int32).reshape(n, k)

    # [n][k/g] each -> interleave (scale, zero) innermost, group-major: [k/g][n][2]
    packed = torch.stack([scales.view(n, -1), zeros.view(n, -1)], dim=2)
    packed = packed.transpose(0, 1).contiguous()
    return codes, packed.to(w_orig.dtype)


def expand_q_groups(x, orig_size, q_group_size):
    """[rows][k/g] -> [rows][k] by repeating each group value g times."""
    rows, k = orig_size
    out = x.reshape(rows, k // q_group_size, 1).expand(rows, k // q_group_size, q_group_size)
    return out.contiguous().view(orig_size)


def extract_scales_and_zeros(scales_and_zeros, w_shape, q_group_size):
    """[k/g][rows][2] -> dense (scales, zeros), each [rows][k]."""
    by_row = scales_and_zeros.transpose(0, 1)
    scales = expand_q_groups(by_row[:, :, 0], w_shape, q_group_size)
    zeros = expand_q_groups(by_row[:, :, 1], w_shape, q_group_size)
    return scales, zeros


def _shared_exponent(block_absmax):
    """floor(log2(absmax)) after rounding the fp32 mantissa up at 1.5 (the reference's
    "even" rounding_mode: add half an exponent step to the bit pattern, keep sign+exp)."""
    bits = block_absmax.to(torch.float32).view(torch.int32)
    bumped = ((bits + (1 << 22)) & (0x1FF << 23)).view(torch.float32)
    tiny = 2.0 ** (-126)
    return torch.floor(torch.log2(bumped + tiny * (bumped == 0).to(bumped.dtype)))


def _round_fp4(v):
    """Round fp32 values to the nearest e2m1 value, ties away from zero, saturating."""
    mag = v.abs()
    e = torch.floor(torch.log2(mag + (v == 0).to(v.dtype))).clamp(min=0)
    step = 2.0**e / 2.0                      # one mantissa bit below the leading one
    q = torch.sign(v) * torch.floor(mag / step + 0.5) * step
    q = q.clamp(-_FP4_MAX, _FP4_MAX)
    q = torch.where(torch.isinf(v), v, q)
    return q


def round_to_mx4(x, q_group_size):
    """Returns (x_q fp32 [n][k] of e2m1 values, exponents fp32 [n][k/g]); the
    reconstruction is x_q * 2**exponent per group."""
    if x.numel() <= 0 or torch.isnan(x).any():
        return x
    x = x.float()
    n, k = x.shape
    assert k % q_group_size == 0, "k must be a multiple of the mx4 group size"
    blocks = x.reshape(n, k // q_group_size, q_group_size)

    exps = _shared_exponent(blocks.abs().amax(dim=-1, keepdim=True))
    blocks = blocks * (exps > -127).to(blocks.dtype)      # flush fp32-subnormal blocks
    exps = exps - _FP4_EMAX
    if (exps > _E8M0_EMAX).any():
        print(f"{exps.max()=} emax={_FP4_EMAX} scale_emax={_E8M0_EMAX} ")
    exps = torch.where(exps > _E8M0_EMAX, torch.full_like(exps, float("nan")), exps)
    exps = exps.clamp(min=-_E8M0_EMAX)

    x_q = _round_fp4(blocks / (2**exps))
    if torch.isnan(x_q * (2**exps)).any():
        raise RuntimeError("NaN encountered while rounding to mx4")
    return x_q.reshape(n, k), exps.reshape(n, k // q_group_size)


def quantize_mx4(x, q_group_size):
    """-> (codes int32 [n][k] in [0,15], e8m0 exponents uint8 [n][k/g])."""
    x_q, x_e = round_to_mx4(x, q_group_size)
    assert x_q.dtype == torch.float32 and x_e.dtype == torch.float32

    q = torch.full(x_q.size(), -128, dtype=torch.int32, device=x_q.device)
    for code, v in enumerate(MX4_VALUES):
        if v == 0.0:
            continue  # +/-0 are told apart by bit pattern below
        q = torch.where(x_q == v, code, q)
    raw = x_q.view(torch.int32)
    q = torch.where(raw == 0, 0, q)
    q = torch.where(raw == -(2**31), 8, q)
    assert (q == -128).sum() == 0, "value not representable in mx4"

    assert (x_e > 128).sum() == 0
    e_int = (x_e + 127).to(torch.uint8)
    return q, e_int


def dequantize_mx4(q, e):
    """codes [n][k] + uint8 exponents [n][k/g] -> fp32 [n][k]."""
    num_groups = e.size(1)
    assert q.size(1) % num_groups == 0 and q.size(1) // num_groups > 0
    g = q.size(1) // num_groups
    table = torch.tensor(MX4_VALUES, dtype=torch.float32, device=q.device)
    vals = table[q.long()].reshape(q.size(0), num_groups, g)
    scale = 2 ** (e.float() - 127).reshape(e.size(0), num_groups, 1)
    return (vals * scale).reshape(q.size())


__all__ = [
    "group_quantize_tensor",
    "expand_q_groups",
    "extract_scales_and_zeros",
    "round_to_mx4",
    "quantize_mx4",
    "dequantize_mx4",
    "MX4_VALUES",
]
