"""Parity cases at the HEADLINE sizes (BASELINE.json configs 1-4): any4 row-wise / global LUT (nf4), int4 and mx4 at
4096^2, 8192^2, 11008^2 and the Llama-3-8B MLP shapes, m in {1, 4, 8, 16}, both weight sides.
TEST INFRASTRUCTURE ONLY (see oracle/__init__.py).

The inputs are synthetic in the PACKED domain (any int32 word is eight valid codes), generated from a CPU generator
seeded by the shape, so this repo's kernels and the unmodified reference extension (oracle/ref_runner.py golden_big)
see bit-identical tensors.  Only a sample of the output columns is kept: tests/golden/golden_big.npz stays small.
"""
import functools

import numpy as np
import torch

SHAPES = [(4096, 4096), (8192, 8192), (11008, 11008), (14336, 4096), (4096, 14336)]  # (n, k)
G = 128
MX_G = 32
N_SAMPLE = 256

NF4 = [-1.0, -0.6962, -0.5251, -0.3949, -0.2844, -0.1848, -0.0911, 0.0, 0.0796, 0.1609, 0.2461, 0.3379, 0.4407, 0.5626,
       0.723, 1.0]  # kmeans.py:17


def cases():
    out = []
    for (n, k) in SHAPES:
        for fmt in ("any4r", "any4g", "int4", "mx4"):
            for m in (1, 4, 8, 16):
                out.append(dict(fmt=fmt, side="right", n=n, k=k, m=m))
        for fmt in ("int4", "any4r"):
            for m in (1, 16):
                out.append(dict(fmt=fmt, side="left", n=n, k=k, m=m))
    # mx4 with e8m0 exponent 0 (2^-127) in some groups: whatever the reference build does with it is the contract
    out.append(dict(fmt="mx4", side="right", n=4096, k=4096, m=1, exp0=True))
    return out


def case_id(c):
    return f"big-{c['fmt']}-{c['side']}-n{c['n']}-k{c['k']}-m{c['m']}" + ("-exp0" if c.get("exp0") else "")


def sample_cols(n):
    return torch.linspace(0, n - 1, N_SAMPLE).round().long().unique()


@functools.lru_cache(maxsize=2)
def shape_inputs(n, k):
    """Everything that depends on the shape only (shared by the formats / m / sides of one shape)."""
    gen = torch.Generator().manual_seed(n * 31 + k)
    words = torch.randint(-2**31, 2**31 - 1, (n * k // 8,), generator=gen, dtype=torch.int64).to(torch.int32)
    lut = ((torch.rand(n, 16, generator=gen) * 15).sort(1).values.bfloat16() - 8)
    sz = torch.stack([torch.rand(k // G, n, generator=gen) * 0.01 + 0.001, torch.randn(k // G, n, generator=gen) * 0.01],
                     dim=2).bfloat16().contiguous()
    exps = torch.randint(118, 130, (n, k // MX_G), generator=gen, dtype=torch.int32).to(torch.uint8)
    x = torch.randn(16, k, generator=gen).bfloat16()
    return dict(words=words, lut=lut, sz=sz, exps=exps, x=x)


def run_ops(c, device="cuda:0"):
    """Evaluate the case through torch.ops.tinygemm.* (whichever library registered it); returns y[:, sample]."""
    ops = torch.ops.tinygemm
    n, k, m = c["n"], c["k"], c["m"]
    s = shape_inputs(n, k)
    right = c["side"] == "right"
    w = s["words"].to(device).view(n // 8, k // 64, 32, 2) if right else s["words"].to(device).view(n // 16, k // 64, 32, 4)
    x = s["x"][:m].contiguous().to(device)
    a, b = (x, w) if right else (w, x)
    fmt = c["fmt"]
    if fmt == "int4":
        y = ops.tinygemm_y_f16RM_x_f16RM_w_int4TC(a, b, G, s["sz"].to(device), right)
    elif fmt == "any4r":
        y = ops.tinygemm_y_f16RM_x_f16RM_w_any4TC(a, b, G, s["sz"].to(device), s["lut"].to(device), right)
    elif fmt == "any4g":
        y = ops.tinygemm_y_f16RM_x_f16RM_w_any4TC(a, b, G, s["sz"].to(device), torch.tensor(NF4).bfloat16().to(device), right)
    elif fmt == "mx4":
        e = s["exps"].clone()
        if c.get("exp0"):
            e[::7, ::3] = 0
        y = ops.tinygemm_y_f16RM_x_f16RM_w_mx4TC(a, b, MX_G, e.to(device), right)
    else:
        raise ValueError(fmt)
    return y[:, sample_cols(n).to(device)].cpu()


# ---- int8 and 16-bit weights at the headline sizes (tests/golden/golden_big_w8.npz) ----
W8_SHAPES = [(4096, 4096), (8192, 8192), (4096, 11008)]


def cases_w8():
    out = []
    for (n, k) in W8_SHAPES:
        for side in ("right", "left"):
            for m in (1, 8, 16):
                out.append(dict(fmt="int8", side=side, n=n, k=k, m=m))
    for m in (1, 16):
        out.append(dict(fmt="f16", side="right", n=4096, k=4096, m=m))
    return out


@functools.lru_cache(maxsize=1)
def shape_inputs_w8(n, k):
    gen = torch.Generator().manual_seed(n * 37 + k + 1)
    words = torch.randint(-2**31, 2**31 - 1, (n * k // 4,), generator=gen, dtype=torch.int64).to(torch.int32)
    sz = torch.stack([torch.rand(k // G, n, generator=gen) * 0.01 + 0.001, torch.randn(k // G, n, generator=gen) * 0.01],
                     dim=2).bfloat16().contiguous()
    x = torch.randn(16, k, generator=gen).bfloat16()
    w16 = torch.randn(n // 8, k // 32, 32, 8, generator=gen).bfloat16() if n * k <= 4096 * 4096 else None
    return dict(words=words, sz=sz, x=x, w16=w16)


def run_ops_w8(c, device="cuda:0"):
    """int8: B layout inner-k 4 [n/8][k/64][32][4], A layout inner-k 2 [n/16][k/32][32][4]; 16-bit: B layout inner-k 2."""
    ops = torch.ops.tinygemm
    n, k, m = c["n"], c["k"], c["m"]
    s = shape_inputs_w8(n, k)
    right = c["side"] == "right"
    x = s["x"][:m].contiguous().to(device)
    if c["fmt"] == "f16":
        y = ops.tinygemm_y_f16RM_x_f16RM_w_f16TC(x, s["w16"].to(device), True)
    else:
        w = s["words"].to(device).view(n // 8, k // 64, 32, 4) if right else s["words"].to(device).view(n // 16, k // 32, 32, 4)
        a, b = (x, w) if right else (w, x)
        y = ops.tinygemm_y_f16RM_x_f16RM_w_int8TC(a, b, G, s["sz"].to(device), right)
    return y[:, sample_cols(n).to(device)].cpu()


def to_u16(t):
    return t.contiguous().view(torch.int16).numpy().view(np.uint16)


def from_u16(a):
    return torch.from_numpy(np.ascontiguousarray(a).view(np.int16)).view(torch.bfloat16)
