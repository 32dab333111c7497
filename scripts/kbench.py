#!/usr/bin/env python
"""Quick graph-timed us/GEMV of the hot kernel for a few shapes (tuning helper; env knobs TG_W4_*)."""
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import tinygemm  # noqa: E402,F401
from bench import G, algorithmic_bytes, synth_layer  # noqa: E402


def main():
    from any4_b200 import functional as tgf
    tgf.set_static_weights(os.environ.get('KB_STATIC', '1') == '1')
    _native_lib = __import__('any4_b200._native', fromlist=['capi']).capi()
    _native_lib.tg_set_option(0, int(os.environ.get('KB_PDL', '1')))
    dev = torch.device("cuda:0")
    op = torch.ops.tinygemm.tinygemm_y_f16RM_x_f16RM_w_any4TC
    out = {}
    checks = {}
    gbps = {}
    for arg in sys.argv[1:] or ["4096", "8192", "11008"]:   # n (square) or NxK
        n, k = (int(v) for v in arg.split("x")) if "x" in arg else (int(arg), int(arg))
        tag = arg
        nbytes = algorithmic_bytes(n, k)
        copies = max(3, int(float(os.environ.get('KB_L2X', '2.6')) * 126e6 / nbytes) + 1)
        layers = [synth_layer(n, k, 10 + i, dev) for i in range(copies)]
        x = torch.randn(int(os.environ.get('KB_M', '1')), k, device=dev).bfloat16()

        side_a = os.environ.get("KB_SIDE", "B") == "A"
        if side_a:  # same bytes read as the A int4 layout (ik = 4), LUT/scales per (padded) row as before
            layers = [(w.view(n // 16, k // 64, 32, 4), lut, sz) for w, lut, sz in layers]

        def step():
            return [op(w, x, G, sz, lut, False) if side_a else op(x, w, G, sz, lut, True) for w, lut, sz in layers]

        for _ in range(3):
            step()
        torch.cuda.synchronize()
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g):
            outs = step()
        g.replay()
        torch.cuda.synchronize()
        # the graph's results (static weights + PDL, back to back) must equal plain stream-ordered launches bit for bit
        tgf.set_static_weights(False)
        _native_lib.tg_set_option(0, 0)
        ok = all(torch.equal(a, b) for a, b in zip(outs, step()))
        tgf.set_static_weights(os.environ.get('KB_STATIC', '1') == '1')
        _native_lib.tg_set_option(0, int(os.environ.get('KB_PDL', '1')))
        checks[tag] = ok
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(10):
            g.replay()
        e1.record()
        torch.cuda.synchronize()
        us = e0.elapsed_time(e1) * 1e3 / (10 * copies)
        out[tag] = round(us, 2)
        gbps[tag] = round(nbytes / us / 1e3)
        del layers, outs, g
        torch.cuda.empty_cache()
    print(json.dumps({"env": {k: v for k, v in os.environ.items() if k.startswith("TG_W4")}, "side": os.environ.get("KB_SIDE", "B"), "m": int(os.environ.get("KB_M", "1")), "l2x": os.environ.get("KB_L2X", "2.6"), "us_per_gemv": out, "bit_equal_to_plain_launches": checks,
                      "GBps": gbps}))


if __name__ == "__main__":
    main()
