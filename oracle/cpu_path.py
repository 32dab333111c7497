"""torch-CPU port of the reference's own CPU path for the any4 Linear forward.
TEST INFRASTRUCTURE / CPU BASELINE ONLY (see oracle/__init__.py).

The reference has no packed CPU kernel: on CPU a quantized Linear is evaluated by
densifying the weight with `quantize.anyq_dequantize_tensor` (quantize.py:612-637 ->
`extract_scales_and_zeros` :151-158, `expand_q_groups` :178-181, `degroup_q` :160-174)
and calling `torch.nn.functional.linear`.  The arithmetic is deliberately the
reference's: gather, `(w - 2^(n_bit-1)) * scales + zeros` as three separately rounded
element-wise ops in the LUT's dtype - NOT the GPU kernel's single-rounded FMA.
`any4` here is the *un-centred* LUT in [0, 15] code space (the GPU module stores
`lut - 8`, quantize.py:893).

Pinned bit-exactly against the imported reference by tests/golden/golden_cpu.npz.
"""
import torch


def expand_q_groups(x, orig_size, q_group_size):
    out = x.reshape(orig_size[0], orig_size[1] // q_group_size, 1)
    out = out.expand(orig_size[0], orig_size[1] // q_group_size, q_group_size)
    return out.contiguous().view(orig_size)


def extract_scales_and_zeros(scales_and_zeros, w_shape, q_group_size):
    t = scales_and_zeros.transpose(0, 1)
    return (
        expand_q_groups(t[:, :, 0], w_shape, q_group_size),
        expand_q_groups(t[:, :, 1], w_shape, q_group_size),
    )


def anyq_dequantize(assign, any4, scales_and_zeros, n_bit=4, q_group_size=128, per_row=True):
    """Dense [n][k] weight in any4.dtype.  Mirrors quantize.py:612-637 with
    new_grouping=False, scale_only=False."""
    orig_shape = assign.shape
    if not per_row:
        assign = assign.reshape(1, -1)
        scales_and_zeros = scales_and_zeros.reshape(-1, 1, 2)
        wc = any4[assign]
    else:
        wc = torch.gather(input=any4, dim=1, index=assign.long())
    scales, zeros = extract_scales_and_zeros(scales_and_zeros, assign.shape, q_group_size)
    wc = wc - (2 ** (n_bit - 1))
    wdeq = wc * scales + zeros
    if not per_row:
        wdeq = wdeq.reshape(orig_shape)
    return wdeq


def any4_linear_forward(x, assign, any4, scales_and_zeros, q_group_size=128, per_row=True, bias=None):
    """One forward of the reference's CPU path: densify, then F.linear."""
    w = anyq_dequantize(assign, any4, scales_and_zeros, 4, q_group_size, per_row)
    return torch.nn.functional.linear(x, w, bias)
