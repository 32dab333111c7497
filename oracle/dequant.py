"""Bit-exact CPU restatement of tinygemm's dequantisation + GEMM numerics.
TEST INFRASTRUCTURE ONLY (see oracle/__init__.py).

Numerics contract restated here (T = activation dtype, bf16 or fp16):
  1. v = LUT[code] as T
       any4 : user LUT, global [16] or per weight row [rows][16]
              (Dequantization.cuh:55-90 row-wise, :92-131 global)
       int4 : exactly code - 8        (Dequantization.cuh:136-178 bf16, :183-260 fp16)
       int8 : exactly code - 128      (Dequantization.cuh:265-328)
       mx4  : T(kMX4_Values[code])    (FloatDefs.cuh:18-34, MatrixLayoutB.cuh:794-799)
  2. int4/any4/int8: w = fma.rn(v, scale, zero) with ONE rounding to T
       (FloatDefs.cuh:87-96 `__hfma2`; MatrixLayoutB.cuh:1042-1046, MatrixLayoutA.cuh:747-754)
     mx4: w = v * T(2^(e-127)), e == 255 -> NaN
       (Dequantization.cuh:331-351, MatrixLayoutB.cuh:1086-1088)
     scale/zero of element (row r, col c) = qScaleAndZeros[c // g][r][0..1];
     exponent = mx4Exponents[r][c // g].
  3. y = RN_T( sum_k x*w ) - the reference accumulates exact bf16xbf16 products in fp32
     inside mma.sync in an implementation-defined order (TinyGemmImpl.cuh:211-216,
     306-340) and rounds once (MatrixLayoutA.cuh:185-186).  The oracle accumulates in
     float64, i.e. it is the value every faithful fp32-accumulating kernel agrees with
     up to fp32 round-off before the final rounding.
"""
import numpy as np
import torch

MX4_VALUES = np.array(
    [0.0, 0.5, 1.0, 1.5, 2.0, 3.0, 4.0, 6.0, -0.0, -0.5, -1.0, -1.5, -2.0, -3.0, -4.0, -6.0],
    dtype=np.float64,
)


# ---------------------------------------------------------------------------
# exact rounding helpers
# ---------------------------------------------------------------------------
def _two_sum(a, b):
    s = a + b
    bb = s - a
    err = (a - (s - bb)) + (b - bb)
    return s, err


def _round_to_odd64(hi, lo):
    """Given exact value hi+lo (hi = RN64(exact)), return the float64 obtained by
    rounding the exact value to odd.  Rounding that to any narrower format with RN
    then equals a single correct rounding of the exact value."""
    hi = np.ascontiguousarray(hi, dtype=np.float64)
    bits = hi.view(np.int64).copy()
    inexact = (lo != 0) & np.isfinite(hi)
    # truncate toward zero: if lo has the opposite sign of hi, step one ulp toward 0
    toward_zero = inexact & (np.signbit(lo) != np.signbit(hi))
    bits = np.where(toward_zero, bits - 1, bits)
    bits = np.where(inexact, bits | 1, bits)
    return bits.view(np.float64)


def _f64_to_bf16_torch(x_ro):
    """float64 (already rounded-to-odd at 53 bits) -> torch.bfloat16 with ONE rounding."""
    bits = np.ascontiguousarray(x_ro).view(np.int64)
    drop_mask = np.int64((1 << 29) - 1)
    dropped = bits & drop_mask
    keep = bits & ~drop_mask
    keep = np.where((dropped != 0) & np.isfinite(x_ro), keep | np.int64(1 << 29), keep)
    f32 = keep.view(np.float64).astype(np.float32)   # exact for the normal fp32 range
    return torch.from_numpy(f32).to(torch.bfloat16)  # RN-even, fp32 -> bf16


def round_f64_pair(hi, lo, dtype):
    """Correctly round the exact value hi+lo to torch dtype (bf16 / fp16)."""
    ro = _round_to_odd64(hi, lo)
    if dtype == torch.bfloat16:
        return _f64_to_bf16_torch(ro)
    if dtype == torch.float16:
        with np.errstate(over="ignore"):
            return torch.from_numpy(ro.astype(np.float16))
    raise TypeError(dtype)


def fma_rn(v, s, z, dtype):
    """fma.rn in `dtype`: v, s, z are torch tensors of that dtype (broadcastable)."""
    v64 = v.to(torch.float64).numpy()
    s64 = s.to(torch.float64).numpy()
    z64 = z.to(torch.float64).numpy()
    p = v64 * s64                      # exact: <= 22 significant bits
    hi, lo = _two_sum(p, np.broadcast_to(z64, p.shape))
    return round_f64_pair(hi, lo, dtype)


# ---------------------------------------------------------------------------
# dequantisation to a dense [rows][k] matrix of dtype T
# ---------------------------------------------------------------------------
def _expand_groups(t, k, g):
    # t: [k/g][rows] -> [rows][k]
    return t.transpose(0, 1).repeat_interleave(g, dim=1)[:, :k]


def dequant_lut(codes, lut, scales_and_zeros, g, dtype):
    """any4 / nf4 / int4-as-LUT.  codes [rows][k] int, lut [16] or [rows][16] (dtype),
    scales_and_zeros [k/g][rows][2] (dtype)."""
    codes = torch.as_tensor(codes).long()
    rows, k = codes.shape
    if lut.dim() == 1:
        v = lut[codes]
    else:
        v = torch.gather(lut, 1, codes)
    s = _expand_groups(scales_and_zeros[:, :, 0], k, g)
    z = _expand_groups(scales_and_zeros[:, :, 1], k, g)
    return fma_rn(v, s, z, dtype)


def dequant_int4(codes, scales_and_zeros, g, dtype):
    lut = (torch.arange(16, dtype=torch.float32) - 8).to(dtype)   # exact in both dtypes
    return dequant_lut(codes, lut, scales_and_zeros, g, dtype)


def dequant_int8(codes, scales_and_zeros, g, dtype):
    codes = torch.as_tensor(codes).long()
    rows, k = codes.shape
    v = (codes - 128).to(torch.float32).to(dtype)                  # exact in both dtypes
    s = _expand_groups(scales_and_zeros[:, :, 0], k, g)
    z = _expand_groups(scales_and_zeros[:, :, 1], k, g)
    return fma_rn(v, s, z, dtype)


def e8m0_to_dtype(e_u8, dtype):
    """uint8 e8m0 -> T: 2^(e-127), 255 -> NaN (Dequantization.cuh:331-351)."""
    e = torch.as_tensor(e_u8).to(torch.int32)
    val = torch.pow(torch.tensor(2.0, dtype=torch.float64), (e - 127).to(torch.float64))
    val = torch.where(e == 255, torch.tensor(float("nan"), dtype=torch.float64), val)
    return val.to(torch.float32).to(dtype)


def dequant_mx4(codes, exponents, g, dtype=torch.bfloat16):
    """codes [rows][k] int, exponents [rows][k/g] uint8."""
    codes = torch.as_tensor(codes).long()
    rows, k = codes.shape
    v = torch.from_numpy(MX4_VALUES)[codes].to(dtype)
    scale = e8m0_to_dtype(exponents, dtype).repeat_interleave(g, dim=1)[:, :k]
    prod = v.to(torch.float64) * scale.to(torch.float64)           # exact
    return prod.to(torch.float32).to(dtype)                        # RN once (fp32 step is exact)


# ---------------------------------------------------------------------------
# GEMM
# ---------------------------------------------------------------------------
def gemm_f64(x, w):
    """x [m][k], w [rows][k] (same dtype T) -> float64 [m][rows], exact products."""
    return x.to(torch.float64) @ w.to(torch.float64).t()


def gemm(x, w):
    """y = RN_T(x @ w^T) with wide accumulation."""
    acc = gemm_f64(x, w)
    if x.dtype == torch.bfloat16:
        # float64 -> bf16 directly: go through round-to-odd-at-fp32 to avoid double rounding
        a = acc.numpy()
        return _f64_to_bf16_torch(a)
    with np.errstate(over="ignore"):
        return torch.from_numpy(acc.numpy().astype(np.float16))


def ulp_distance(a, b):
    """Element-wise distance in units of representable values for bf16/fp16 tensors
    (both finite).  0 = bit-equal up to the sign of zero."""
    def key(t):
        i = t.view(torch.int16).to(torch.int32)
        return torch.where(i < 0, -(i & 0x7FFF), i)
    return (key(a) - key(b)).abs()
