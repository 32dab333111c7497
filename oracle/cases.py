"""Seeded parity cases shared by the GPU tests, the golden-vector generator and smoke().
TEST INFRASTRUCTURE ONLY (see oracle/__init__.py).

A case is a plain dict.  `make_inputs(case)` builds its CPU tensors deterministically from
`case["seed"]`; `run_ops(case, inputs, device)` evaluates it through `torch.ops.tinygemm.*`
(whichever library registered that namespace in the current process: this repo's, or the
unmodified reference extension inside oracle/ref_runner.py); `oracle_output(case, inputs)`
evaluates it with the CPU restatement (oracle/layouts.py, oracle/dequant.py).

Input recipes follow SURVEY.md 8(d): codes = randint(0, 16), LUT = sorted rand*15 - 8 in the
activation dtype (row-wise or global), scale = rand*0.01 + 0.001, zero = randn*0.01; int4 / int8
from group_quantize_tensor(randn), mx4 from quantize_mx4(randn).
"""
import numpy as np
import torch

from . import dequant, layouts

DT = {"bf16": torch.bfloat16, "fp16": torch.float16}


def _u16(t):
    return t.contiguous().view(torch.int16).numpy().view(np.uint16)


def _from_u16(a, dtype):
    return torch.from_numpy(np.ascontiguousarray(a).view(np.int16)).view(dtype)


# ------------------------------------------------------------------------------------------
# case lists
# ------------------------------------------------------------------------------------------
def convert_cases():
    cases = []
    seed = 1000
    for dt in ("bf16", "fp16"):
        for (m, k) in [(1, 1), (5, 7), (16, 16), (17, 33), (15, 15), (32, 64), (48, 80), (3, 130)]:
            cases.append(dict(kind="convert", op="A", dt=dt, rows=m, k=k, ik=1, seed=seed)); seed += 1
            for ik in (1, 2):
                cases.append(dict(kind="convert", op="B", dt=dt, rows=m, k=k, ik=ik, seed=seed)); seed += 1
    for (m, k) in [(16, 64), (5, 64), (33, 128), (64, 256), (19, 192), (8, 320)]:
        for ik in (1, 2, 4):
            cases.append(dict(kind="convert", op="Aint4", rows=m, k=k, ik=ik, seed=seed)); seed += 1
        for ik in (1, 2):
            cases.append(dict(kind="convert", op="Aint8", rows=m, k=k, ik=ik, seed=seed)); seed += 1
        for ik in (2, 4, 8):
            if k % (ik * 16) == 0:
                cases.append(dict(kind="convert", op="Bint4", rows=m, k=k, ik=ik, seed=seed)); seed += 1
        for ik in (1, 2, 4):
            if k % (ik * 16) == 0:
                cases.append(dict(kind="convert", op="Bint8", rows=m, k=k, ik=ik, seed=seed)); seed += 1
    # ragged k for the A-side packers (zero padding inside the last tile)
    for (m, k) in [(7, 40), (20, 100)]:
        for ik in (1, 2, 4):
            cases.append(dict(kind="convert", op="Aint4", rows=m, k=k, ik=ik, seed=seed)); seed += 1
    return cases


def gemm_cases(full=True):
    """kind="gemm": fmt in {int4, any4g, any4r, mx4, int8, f16}; side "right" (B layout) / "left"
    (A layout); api "RM" / "TC"; m = activation rows; n = true weight rows; k; g; ik."""
    cases = []
    seed = 5000

    def add(**kw):
        nonlocal seed
        kw.setdefault("x_ik", 1)
        cases.append(dict(kind="gemm", seed=seed, **kw))
        seed += 1

    # --- the hot path: B-layout 4-bit, row-major activations ---
    for fmt in ("any4r", "any4g", "int4", "mx4"):
        for dt in ("bf16", "fp16"):
            if fmt == "mx4" and dt == "fp16":
                continue
            for ik in (2, 4, 8):
                for g in (32, 64, 128, 256):
                    if fmt == "mx4" and g != 32 and ik != 4:
                        continue
                    add(fmt=fmt, dt=dt, side="right", api="RM", m=1, n=64, k=512, g=g, ik=ik)
            for m in (2, 3, 4, 5, 8, 16, 19):
                add(fmt=fmt, dt=dt, side="right", api="RM", m=m, n=64, k=256, g=64, ik=4)
            # ragged rows (n pads to a multiple of 8, not of 32), k not a multiple of 128
            add(fmt=fmt, dt=dt, side="right", api="RM", m=1, n=40, k=320, g=32, ik=2)
            add(fmt=fmt, dt=dt, side="right", api="RM", m=3, n=20, k=192, g=64, ik=4)
            # split-k cluster path (few rows, long k) and a multi-wave shape
            add(fmt=fmt, dt=dt, side="right", api="RM", m=1, n=64, k=4096, g=128, ik=4)
            add(fmt=fmt, dt=dt, side="right", api="RM", m=2, n=32, k=8192, g=128, ik=8)
    if full:
        add(fmt="any4r", dt="bf16", side="right", api="RM", m=1, n=1024, k=1024, g=128, ik=4)
        add(fmt="any4r", dt="bf16", side="right", api="RM", m=4, n=512, k=2048, g=128, ik=4)

    # --- A-layout 4-bit (Int4Linear default), row-major ---
    for fmt in ("int4", "any4r", "any4g", "mx4"):
        for dt in ("bf16", "fp16"):
            if fmt == "mx4" and dt == "fp16":
                continue
            for ik in (1, 2, 4):
                add(fmt=fmt, dt=dt, side="left", api="RM", m=1, n=64, k=256, g=32 if fmt == "mx4" else 64, ik=ik)
            for m in (5, 16, 19):
                add(fmt=fmt, dt=dt, side="left", api="RM", m=m, n=48, k=512, g=128 if fmt != "mx4" else 32, ik=4)
            add(fmt=fmt, dt=dt, side="left", api="RM", m=2, n=20, k=256, g=32, ik=2)

    # --- int8 ---
    for dt in ("bf16", "fp16"):
        for ik in (1, 2, 4):
            add(fmt="int8", dt=dt, side="right", api="RM", m=3, n=40, k=256, g=64, ik=ik)
        for ik in (1, 2):
            add(fmt="int8", dt=dt, side="left", api="RM", m=5, n=48, k=256, g=128, ik=ik)
        add(fmt="int8", dt=dt, side="right", api="RM", m=1, n=64, k=1024, g=32, ik=4)

    # --- 16-bit weights ---
    for dt in ("bf16", "fp16"):
        for ik in (1, 2):
            add(fmt="f16", dt=dt, side="right", api="RM", m=3, n=40, k=256, g=0, ik=ik)
        add(fmt="f16", dt=dt, side="left", api="RM", m=5, n=48, k=512, g=0, ik=1)

    # --- tensor-core-layout activations / outputs ---
    for fmt in ("int4", "any4r", "any4g", "mx4", "int8", "f16"):
        for dt in ("bf16", "fp16"):
            if fmt == "mx4" and dt == "fp16":
                continue
            g = 0 if fmt == "f16" else 32
            add(fmt=fmt, dt=dt, side="right", api="TC", m=5, n=40, k=256, g=g, ik=2 if fmt != "f16" else 1)
            add(fmt=fmt, dt=dt, side="right", api="TC", m=19, n=64, k=512, g=g, ik=4 if fmt not in ("f16",) else 2)
            for x_ik in (1, 2):
                add(fmt=fmt, dt=dt, side="left", api="TC", m=19, n=48, k=256, g=g, ik=2 if fmt != "f16" else 1, x_ik=x_ik)
    return cases


def all_cases():
    return convert_cases() + gemm_cases()


def case_id(c):
    if c["kind"] == "convert":
        return f"cv-{c['op']}-{c.get('dt', 'i32')}-{c['rows']}x{c['k']}-ik{c['ik']}"
    return (f"mm-{c['fmt']}-{c['dt']}-{c['side']}-{c['api']}-m{c['m']}-n{c['n']}-k{c['k']}-g{c['g']}-ik{c['ik']}"
            f"-x{c['x_ik']}")


# ------------------------------------------------------------------------------------------
# inputs
# ------------------------------------------------------------------------------------------
def _pad_rows(c):
    t = 8 if c["side"] == "right" else 16
    return (c["n"] + t - 1) // t * t


def make_inputs(c):
    gen = torch.Generator().manual_seed(c["seed"])
    if c["kind"] == "convert":
        if c["op"] in ("A", "B"):
            x = torch.randn(c["rows"], c["k"], generator=gen).to(DT[c["dt"]])
            return dict(x=x)
        hi = 16 if c["op"].endswith("int4") else 256
        return dict(codes=torch.randint(0, hi, (c["rows"], c["k"]), generator=gen, dtype=torch.int32))

    from any4_b200 import utils as host  # host-side quantizers (pinned against the reference in tests)

    dt = DT[c["dt"]]
    m, n, k, g = c["m"], c["n"], c["k"], c["g"]
    rows = _pad_rows(c)
    x = torch.randn(m, k, generator=gen).to(dt)
    out = dict(x=x)
    fmt = c["fmt"]
    if fmt in ("any4r", "any4g"):
        out["codes"] = torch.randint(0, 16, (n, k), generator=gen, dtype=torch.int32)
        nl = rows if fmt == "any4r" else 1
        raw = (torch.rand(nl, 16, generator=gen) * 15).sort(1).values.to(dt)
        lut = raw - 8  # computed in dt, as quantize.py:893 does
        out["lut"] = lut if fmt == "any4r" else lut[0].contiguous()
        scale = torch.rand(k // g, rows, generator=gen) * 0.01 + 0.001
        zero = torch.randn(k // g, rows, generator=gen) * 0.01
        out["sz"] = torch.stack([scale, zero], dim=2).to(dt)
    elif fmt in ("int4", "int8"):
        w = torch.randn(rows, k, generator=gen).to(dt)
        codes, sz = host.group_quantize_tensor(w, 4 if fmt == "int4" else 8, g)
        out["codes"] = codes[:n].contiguous()
        out["sz"] = sz
    elif fmt == "mx4":
        w = torch.randn(rows, k, generator=gen).to(dt)
        codes, e = host.quantize_mx4(w, g)
        out["codes"] = codes[:n].contiguous()
        out["exps"] = e
    elif fmt == "f16":
        out["w"] = (torch.randn(n, k, generator=gen) * 0.1).to(dt)
    else:
        raise ValueError(fmt)
    return out


# ------------------------------------------------------------------------------------------
# evaluation through torch.ops.tinygemm (ours or the reference's)
# ------------------------------------------------------------------------------------------
def run_ops(c, inp, device="cuda:0"):
    ops = torch.ops.tinygemm
    d = {k: v.to(device) for k, v in inp.items()}
    if c["kind"] == "convert":
        op, ik = c["op"], c["ik"]
        if op == "A":
            packed = ops.convert_matrix_to_m16n8k16_A_layout(d["x"], 1)
            back = ops.convert_matrix_from_m16n8k16_A_layout(packed, c["rows"], c["k"])
            return dict(packed=packed.cpu(), back=back.cpu())
        if op == "B":
            packed = ops.convert_matrix_to_m16n8k16_B_layout(d["x"], ik)
            back = ops.convert_matrix_from_m16n8k16_B_layout(packed, c["rows"], c["k"])
            return dict(packed=packed.cpu(), back=back.cpu())
        fn = getattr(ops, f"convert_matrix_to_m16n8k16_{op}_layout")
        return dict(packed=fn(d["codes"], ik).cpu())

    fmt, right, ik = c["fmt"], c["side"] == "right", c["ik"]
    if fmt == "f16":
        w2 = (ops.convert_matrix_to_m16n8k16_B_layout(d["w"], ik) if right
              else ops.convert_matrix_to_m16n8k16_A_layout(d["w"], 1))
    else:
        bits = "int8" if fmt == "int8" else "int4"
        w2 = getattr(ops, f"convert_matrix_to_m16n8k16_{'B' if right else 'A'}{bits}_layout")(d["codes"], ik)
    x = d["x"]
    tc = c["api"] == "TC"
    if tc:
        x2 = (ops.convert_matrix_to_m16n8k16_A_layout(x, 1) if right
              else ops.convert_matrix_to_m16n8k16_B_layout(x, c["x_ik"]))
    else:
        x2 = x
    A, B = (x2, w2) if right else (w2, x2)
    L = "TC" if tc else "RM"
    if fmt == "int4":
        y = getattr(ops, f"tinygemm_y_f16{L}_x_f16{L}_w_int4TC")(A, B, c["g"], d["sz"], right)
    elif fmt in ("any4r", "any4g"):
        y = getattr(ops, f"tinygemm_y_f16{L}_x_f16{L}_w_any4TC")(A, B, c["g"], d["sz"], d["lut"], right)
    elif fmt == "mx4":
        y = getattr(ops, f"tinygemm_y_f16{L}_x_f16{L}_w_mx4TC")(A, B, c["g"], d["exps"], right)
    elif fmt == "int8":
        y = getattr(ops, f"tinygemm_y_f16{L}_x_f16{L}_w_int8TC")(A, B, c["g"], d["sz"], right)
    else:
        y = getattr(ops, f"tinygemm_y_f16{L}_x_f16{L}_w_f16TC")(A, B, right)
    if tc:
        y = (ops.convert_matrix_from_m16n8k16_A_layout(y, c["m"], c["n"]) if right
             else ops.convert_matrix_from_m16n8k16_B_layout(y, c["m"], c["n"]))
    else:
        y = y[:, : c["n"]]
    return dict(y=y.contiguous().cpu())


# ------------------------------------------------------------------------------------------
# CPU oracle
# ------------------------------------------------------------------------------------------
def oracle_output(c, inp):
    if c["kind"] == "convert":
        op, ik = c["op"], c["ik"]
        if op == "A":
            return dict(packed=_from_u16(layouts.to_A(_u16(inp["x"])), inp["x"].dtype), back=inp["x"])
        if op == "B":
            return dict(packed=_from_u16(layouts.to_B(_u16(inp["x"]), ik), inp["x"].dtype), back=inp["x"])
        fn = getattr(layouts, f"to_{op}")
        return dict(packed=torch.from_numpy(fn(inp["codes"].numpy(), ik)))

    dt = DT[c["dt"]]
    fmt, n, g = c["fmt"], c["n"], c["g"]
    if fmt == "f16":
        w = inp["w"]
    else:
        codes = inp["codes"]
        if fmt == "int4":
            w = dequant.dequant_int4(codes, inp["sz"][:, :n], g, dt)
        elif fmt == "int8":
            w = dequant.dequant_int8(codes, inp["sz"][:, :n], g, dt)
        elif fmt == "any4g":
            w = dequant.dequant_lut(codes, inp["lut"], inp["sz"][:, :n], g, dt)
        elif fmt == "any4r":
            w = dequant.dequant_lut(codes, inp["lut"][:n], inp["sz"][:, :n], g, dt)
        else:
            w = dequant.dequant_mx4(codes, inp["exps"][:n], g, dt)
    return dict(y=dequant.gemm(inp["x"], w), y64=dequant.gemm_f64(inp["x"], w))


def compare_gemm(y, ref, y64=None):
    """Parity metrics between a kernel output and a reference (same dtype): relative max error
    (max|y - ref| / max|ref|), fraction of bit-equal elements, max ulp distance."""
    yf, rf = y.double(), ref.double()
    denom = max(rf.abs().max().item(), 1e-30)
    rel = (yf - rf).abs().max().item() / denom
    ulp = dequant.ulp_distance(y, ref)
    return dict(rel=rel, frac_equal=(ulp == 0).double().mean().item(), max_ulp=int(ulp.max().item()))
