#!/usr/bin/env python
"""In-kernel phase timeline of the hot GEMV (debug tool; uses the tg_debug_set_trace hook).
Needs the library built with tracing compiled in:  TG_NVCC_EXTRA=-DTG_W4_TRACE python -m any4_b200.build -f
(and a plain `python -m any4_b200.build -f` afterwards).
Prints, per phase, the median / p10 / p90 over CTAs of the time since the FIRST CTA started."""
import ctypes
import json
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from any4_b200 import _native  # noqa: E402
import tinygemm  # noqa: E402,F401
from bench import synth_layer, G  # noqa: E402

NAMES = ["entry", "loads_issued", "tma_start", "tma_refill0", "staged", "cons_sync", "full0", "full1", "full2", "full3",
         "loop_end", "cta_sync", "exit"]


def main():
    lib = _native.capi()
    lib.tg_debug_set_trace.argtypes = [ctypes.c_void_p]
    dev = torch.device("cuda:0")
    out = {}
    for n in [int(a) for a in sys.argv[1:]] or [4096, 8192]:
        k = n
        layers = [synth_layer(n, k, 10 + i, dev) for i in range(max(3, int(300e6 / (n * k / 2))))]
        x = torch.randn(1, k, device=dev).bfloat16()
        op = torch.ops.tinygemm.tinygemm_y_f16RM_x_f16RM_w_any4TC
        for w, lut, sz in layers:
            op(x, w, G, sz, lut, True)
        torch.cuda.synchronize()
        ctas = (n // 32)
        buf = torch.zeros(ctas * 16, dtype=torch.int64, device=dev)
        lib.tg_debug_set_trace(ctypes.c_void_p(buf.data_ptr()))
        w, lut, sz = layers[0]
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        op(x, w, G, sz, lut, True)
        e1.record()
        torch.cuda.synchronize()
        lib.tg_debug_set_trace(None)
        t = buf.cpu().numpy().reshape(ctas, 16).astype(np.int64)
        t = t[t[:, 0] > 0]  # persistent grid: only the first gridDim.x slots are used
        ctas = len(t)
        t0 = t[:, 0].min()
        rel = (t[:, :13] - t0) / 1e3
        rel[t[:, :13] == 0] = np.nan
        print(f"== n=k={n}: {ctas} CTAs, event time {e0.elapsed_time(e1) * 1e3:.1f} us, SMs used {len(set(t[:, 15]))}")
        print(f"{'phase':14s} {'p10':>8s} {'median':>8s} {'p90':>8s} {'max':>8s}   (us since first CTA entry)")
        res = {}
        for i, nm in enumerate(NAMES):
            col = rel[:, i]
            if np.all(np.isnan(col)):
                continue
            q = np.nanpercentile(col, [10, 50, 90, 100])
            res[nm] = [round(float(v), 2) for v in q]
            print(f"{nm:14s} {q[0]:8.2f} {q[1]:8.2f} {q[2]:8.2f} {q[3]:8.2f}")
        dur = (t[:, 12] - t[:, 0]) / 1e3
        print(f"CTA lifetime: median {np.median(dur):.2f} us, min {dur.min():.2f}, max {dur.max():.2f}")
        res["cta_lifetime_us"] = [float(np.median(dur)), float(dur.min()), float(dur.max())]
        out[str(n)] = res
    os.makedirs("gpurun_out", exist_ok=True)
    json.dump(out, open("gpurun_out/trace.json", "w"), indent=1)


if __name__ == "__main__":
    main()
