"""Tolerances of the GEMM parity tests (floating point; stated here once).

The kernel contract: dequantised weights bit-identical to the reference, exact
products, fp32 accumulation in an implementation-defined order, ONE round-to-nearest at the end.
So against the exact float64 result `y64` of the same dequantised operands a faithful kernel obeys

    |y - y64| <= ulp_T(y64) / 2  +  ACC_REL * sum_k |x_k * w_k|

where the second term bounds the fp32 accumulation error (tensor-core accumulation is not IEEE
RN; 2^-16 relative to the absolute dot product is > 10x the statistical error at k = 8192).
The north-star tolerance "1e-3 relative bf16" is checked as a relative Frobenius error against
the oracle's correctly rounded output.
"""
import torch

ACC_REL = 2.0 ** -16
FROB_REL = 1e-3
MANT = {torch.bfloat16: 8, torch.float16: 11}


def half_ulp(y64, dtype):
    """half a unit in the last place of dtype at |y64| (normal range)"""
    a = y64.abs().clamp_min(2.0 ** -126 if dtype == torch.bfloat16 else 2.0 ** -14)
    e = torch.floor(torch.log2(a))
    return torch.pow(torch.tensor(2.0, dtype=torch.float64), e - MANT[dtype] + 1) / 2


def check_faithful(y, y64, absdot, dtype, extra_ulps=0.0):
    err = (y.double() - y64).abs()
    tol = half_ulp(y64, dtype) * (1.0 + 2 * extra_ulps) * (1 + 1e-9) + ACC_REL * absdot
    bad = err > tol
    worst = (err / tol).max().item() if err.numel() else 0.0
    return int(bad.sum().item()), worst


def frob_rel(y, ref):
    d = (y.double() - ref.double()).norm().item()
    return d / max(ref.double().norm().item(), 1e-30)
