#!/usr/bin/env python
"""Generate tests/golden/golden_cpu.npz by IMPORTING THE REFERENCE (build container only).

Runs the reference's own Python for the host-side pieces of the tinygemm path on seeded
inputs and stores inputs + outputs:
  * tinygemm_lib.utils.group_quantize_tensor / quantize_mx4 / dequantize_mx4
  * quantize.anyq_dequantize_tensor (+ F.linear): the reference's CPU path
bf16 tensors are stored as their uint16 bit patterns (key suffix `__bf16`).
`/root/reference` does not exist on the GPU box; only the .npz travels.
"""
import os
import sys
import types

import numpy as np
import torch

REF = os.environ.get("ANY4_REFERENCE", "/root/reference")
sys.path.insert(0, REF)
sys.modules.setdefault("bitsandbytes", types.ModuleType("bitsandbytes"))  # unused on this path

import quantize as refq  # noqa: E402
from tinygemm_lib import utils as refu  # noqa: E402

out = {}


def put(name, t):
    if isinstance(t, torch.Tensor):
        if t.dtype == torch.bfloat16:
            out[name + "__bf16"] = t.contiguous().view(torch.int16).numpy().view(np.uint16)
            return
        t = t.contiguous().numpy()
    out[name] = np.asarray(t)


torch.manual_seed(1234)

# --- group_quantize_tensor ---------------------------------------------------
case = 0
for dt in (torch.bfloat16, torch.float16, torch.float32):
    for n_bit in (4, 8):
        for g in (32, 128):
            w = torch.randn(24, 256).to(dt)
            codes, sz = refu.group_quantize_tensor(w, n_bit, g)
            put(f"gq{case}_w", w)
            put(f"gq{case}_codes", codes)
            put(f"gq{case}_sz", sz)
            out[f"gq{case}_meta"] = np.array([n_bit, g])
            case += 1
out["gq_cases"] = np.array(case)

# identity weight (the exactness fixture of the reference tests)
w = torch.eye(128, dtype=torch.bfloat16)
codes, sz = refu.group_quantize_tensor(w, 4, 32)
put("gq_eye_codes", codes)
put("gq_eye_sz", sz)

# --- mx4 ----------------------------------------------------------------------
case = 0
for scale in (1.0, 1e-3, 37.5, 2.0**-120, 2.0**100):
    x = (torch.randn(16, 128) * scale).to(torch.bfloat16)
    if case == 1:
        x[3, 32:64] = 0  # an all-zero block
    q, e = refu.quantize_mx4(x, 32)
    d = refu.dequantize_mx4(q, e)
    put(f"mx{case}_x", x)
    put(f"mx{case}_q", q)
    put(f"mx{case}_e", e)
    put(f"mx{case}_d", d)
    case += 1
out["mx_cases"] = np.array(case)

# --- reference CPU path: anyq_dequantize_tensor + linear -----------------------
case = 0
for per_row in (True, False):
    for g in (64, 128):
        n, k, m = 16, 256, 3
        assign = torch.randint(0, 16, (n, k), dtype=torch.int32)
        any4 = (torch.rand(n, 16) * 15).sort(1).values.to(torch.bfloat16)
        if not per_row:
            any4 = any4[0].contiguous()
        sz = torch.stack(
            [torch.rand(k // g, n) * 0.01 + 0.001, torch.randn(k // g, n) * 0.01], dim=2
        ).to(torch.bfloat16)
        x = torch.randn(m, k).to(torch.bfloat16)
        w = refq.anyq_dequantize_tensor(assign, any4, sz, n_bit=4, q_group_size=g, per_row=per_row)
        y = torch.nn.functional.linear(x, w)
        put(f"cpu{case}_assign", assign)
        put(f"cpu{case}_any4", any4)
        put(f"cpu{case}_sz", sz)
        put(f"cpu{case}_x", x)
        put(f"cpu{case}_w", w)
        put(f"cpu{case}_y", y)
        out[f"cpu{case}_meta"] = np.array([int(per_row), g])
        case += 1
out["cpu_cases"] = np.array(case)

dst = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden_cpu.npz")
np.savez_compressed(dst, **out)
print("wrote", dst, os.path.getsize(dst), "bytes,", len(out), "arrays")
