#!/usr/bin/env python
"""Run the UNMODIFIED reference tinygemm extension (oracle/_ref/tinygemm.so) on the GPU.
TEST INFRASTRUCTURE ONLY.  Must be its own process: the reference registers the same
`torch.ops.tinygemm` namespace as this repo's library.

  python oracle/ref_runner.py golden <out.npz>      evaluate every parity case (oracle/cases.py)
  python oracle/ref_runner.py golden_big <out.npz>  the headline-size cases (oracle/big_cases.py), sampled columns
  python oracle/ref_runner.py golden_big_w8 <out.npz>  the same for int8 / 16-bit weights (cases_w8)
  python oracle/ref_runner.py bench <n> <k> <iters> time the reference any4 GEMV (rotating copies)
  python oracle/ref_runner.py sweep                 the format x m sweep of bench.py (4096^2) on the reference kernels
"""
import json
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
SO = os.path.join(HERE, "_ref", "tinygemm.so")


def load_reference():
    if not os.path.exists(SO):
        raise SystemExit(f"{SO} missing: build it in the build container with `python oracle/build_ref.py`")
    torch.ops.load_library(SO)


def golden(out_path):
    from oracle import cases as C

    load_reference()
    out = {}
    for c in C.all_cases():
        inp = C.make_inputs(c)
        try:
            res = C.run_ops(c, inp, "cuda:0")
        except RuntimeError as e:
            print(f"[ref] {C.case_id(c)} rejected by the reference: {str(e).splitlines()[0][:120]}")
            continue
        for key, t in res.items():
            name = f"{C.case_id(c)}/{key}"
            if t.dtype in (torch.bfloat16, torch.float16):
                out[name] = t.contiguous().view(torch.int16).numpy().view(np.uint16)
            else:
                out[name] = t.numpy()
    torch.cuda.synchronize()
    np.savez_compressed(out_path, **out)
    print(f"[ref] wrote {out_path}: {len(out)} arrays")


def golden_big(out_path):
    """The headline-size cases (oracle/big_cases.py): sampled output columns of the reference kernels."""
    from oracle import big_cases as B

    load_reference()
    out = {}
    for c in B.cases():
        try:
            y = B.run_ops(c, "cuda:0")
        except RuntimeError as e:
            print(f"[ref] {B.case_id(c)} rejected by the reference: {str(e).splitlines()[0][:120]}")
            continue
        out[B.case_id(c)] = B.to_u16(y)
    np.savez_compressed(out_path, **out)
    print(f"[ref] wrote {out_path}: {len(out)} arrays")


def golden_big_w8(out_path):
    """int8 / 16-bit weights at the headline sizes (oracle/big_cases.py cases_w8): sampled output columns."""
    from oracle import big_cases as B

    load_reference()
    out = {}
    for c in B.cases_w8():
        try:
            y = B.run_ops_w8(c, "cuda:0")
        except RuntimeError as e:
            print(f"[ref] {B.case_id(c)} rejected by the reference: {str(e).splitlines()[0][:120]}")
            continue
        out[B.case_id(c)] = B.to_u16(y)
    np.savez_compressed(out_path, **out)
    print(f"[ref] wrote {out_path}: {len(out)} arrays")


def bench(n, k, iters):
    load_reference()
    g = 128
    dev = torch.device("cuda:0")
    nbytes = n * k // 2 + (k // g) * n * 4 + n * 32 + 2 * k + 2 * n
    copies = max(3, int(2.6 * 126e6 / nbytes) + 1)
    gen = torch.Generator(device=dev).manual_seed(0)
    layers = []
    for _ in range(copies):
        w = torch.randint(-2**31, 2**31 - 1, (n // 8, k // 64, 32, 2), generator=gen, device=dev, dtype=torch.int64).to(torch.int32)
        lut = ((torch.rand(n, 16, generator=gen, device=dev) * 15).sort(1).values.bfloat16() - 8)
        sz = torch.stack([torch.rand(k // g, n, generator=gen, device=dev) * 0.01 + 0.001,
                          torch.randn(k // g, n, generator=gen, device=dev) * 0.01], dim=2).bfloat16().contiguous()
        layers.append((w, lut, sz))
    x = torch.randn(1, k, device=dev).bfloat16()
    op = torch.ops.tinygemm.tinygemm_y_f16RM_x_f16RM_w_any4TC
    def step():
        for w, lut, sz in layers:
            op(x, w, g, sz, lut, True)

    for _ in range(3):
        step()
    torch.cuda.synchronize()
    graph = torch.cuda.CUDAGraph()  # same protocol as bench.py: one graph per step, replayed
    with torch.cuda.graph(graph):
        step()
    graph.replay()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        graph.replay()
    e1.record()
    torch.cuda.synchronize()
    us = e0.elapsed_time(e1) * 1e3 / (iters * copies)
    print(json.dumps({"impl": "reference tinygemm (mma.sync, recompiled for sm_100a)", "n": n, "k": k,
                      "us_per_gemv": us, "GBps": nbytes / us / 1e3, "copies": copies}))


def sweep():
    """Same protocol, shapes and op calls as bench.py's format_sweep(), on the reference's kernels."""
    load_reference()
    ops = torch.ops.tinygemm
    n = k = 4096
    g = 128
    dev = torch.device("cuda:0")
    nbytes = n * k // 2 + (k // g) * n * 4 + n * 32 + 2 * k + 2 * n
    copies = max(3, int(2.6 * 126e6 / nbytes) + 1)
    gen = torch.Generator(device=dev).manual_seed(99)
    ws = []
    for _ in range(copies):
        w = torch.randint(-2**31, 2**31 - 1, (n // 8, k // 64, 32, 2), generator=gen, device=dev, dtype=torch.int64).to(torch.int32)
        lut = ((torch.rand(n, 16, generator=gen, device=dev) * 15).sort(1).values.bfloat16() - 8)
        sz = torch.stack([torch.rand(k // g, n, generator=gen, device=dev) * 0.01 + 0.001,
                          torch.randn(k // g, n, generator=gen, device=dev) * 0.01], dim=2).bfloat16().contiguous()
        ws.append((w, lut, sz))
    nf4 = torch.tensor([-1.0, -0.6962, -0.5251, -0.3949, -0.2844, -0.1848, -0.0911, 0.0, 0.0796, 0.1609, 0.2461,
                        0.3379, 0.4407, 0.5626, 0.723, 1.0], device=dev).bfloat16()
    exps = torch.randint(118, 130, (n, k // 32), generator=gen, device=dev, dtype=torch.int32).to(torch.uint8)
    state = {"copies": copies}

    def timed(fn):
        for _ in range(2):
            fn()
        torch.cuda.synchronize()
        gr = torch.cuda.CUDAGraph()
        with torch.cuda.graph(gr):
            fn()
        gr.replay()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(5):
            gr.replay()
        e1.record()
        torch.cuda.synchronize()
        return round(e0.elapsed_time(e1) * 1e3 / (5 * state["copies"]), 2)

    out = {}
    for m in (1, 4, 8, 16):
        x = torch.randn(m, k, device=dev).bfloat16()
        out[f"m{m}"] = {
            "any4_rowwise_g128": timed(lambda: [ops.tinygemm_y_f16RM_x_f16RM_w_any4TC(x, w, g, sz, lut, True) for w, lut, sz in ws]),
            "nf4_global_g128": timed(lambda: [ops.tinygemm_y_f16RM_x_f16RM_w_any4TC(x, w, g, sz, nf4, True) for w, lut, sz in ws]),
            "int4_g128": timed(lambda: [ops.tinygemm_y_f16RM_x_f16RM_w_int4TC(x, w, g, sz, True) for w, lut, sz in ws]),
            "mx4_g32": timed(lambda: [ops.tinygemm_y_f16RM_x_f16RM_w_mx4TC(x, w, 32, exps, True) for w, lut, sz in ws]),
        }
    x = torch.randn(1, k, device=dev).bfloat16()
    wa = [w.view(n // 16, k // 64, 32, 4) for w, _, _ in ws]
    out["m1"]["int4_g128_A_layout(Int4Linear default)"] = timed(
        lambda: [ops.tinygemm_y_f16RM_x_f16RM_w_int4TC(w, x, g, sz, False) for w, (_, _, sz) in zip(wa, ws)])
    w8 = [torch.randint(-2**31, 2**31 - 1, (n // 8, k // 64, 32, 4), generator=gen, device=dev, dtype=torch.int64).to(torch.int32)
          for _ in range(16)]
    w16 = [torch.randn(n // 8, k // 32, 32, 8, generator=gen, device=dev).bfloat16() for _ in range(8)]
    sz0 = ws[0][2]
    state["copies"] = 16
    out["m1"]["int8_g128_B_layout"] = timed(lambda: [ops.tinygemm_y_f16RM_x_f16RM_w_int8TC(x, w, g, sz0, True) for w in w8])
    state["copies"] = 8
    out["m1"]["bf16_weights_B_layout"] = timed(lambda: [ops.tinygemm_y_f16RM_x_f16RM_w_f16TC(x, w, True) for w in w16])
    out["unit"] = "us per GEMM at n=k=4096, bf16, CUDA graph, weights rotated through > 2.5x L2"
    print(json.dumps({"impl": "reference tinygemm (mma.sync, recompiled for sm_100a)", "format_sweep_us": out}))


if __name__ == "__main__":
    if sys.argv[1] == "sweep":
        sweep()
    elif sys.argv[1] == "golden":
        golden(sys.argv[2])
    elif sys.argv[1] == "golden_big":
        golden_big(sys.argv[2])
    elif sys.argv[1] == "golden_big_w8":
        golden_big_w8(sys.argv[2])
    elif sys.argv[1] == "bench":
        bench(int(sys.argv[2]), int(sys.argv[3]), int(sys.argv[4]))
