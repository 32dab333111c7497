#!/usr/bin/env python
"""In-kernel phase timeline of the tcgen05 GEMV (debug tool; uses the tg_debug_set_trace hook).
Needs the library built with tracing compiled in:  TG_NVCC_EXTRA=-DTG_W4_TRACE python -m any4_b200.build -f
(and a plain `python -m any4_b200.build -f` afterwards).
Prints, per phase, the median / p10 / p90 over CTAs of the time since the FIRST CTA started.
  TR_M=<rows>  TR_STATIC=0/1  TR_PREV=1 (launch a second GEMV right before, to see the PDL overlap)"""
import ctypes
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from any4_b200 import _native  # noqa: E402
import tinygemm  # noqa: E402,F401
from bench import synth_layer, G  # noqa: E402

NAMES = {0: "entry", 1: "setup_done", 2: "table0", 11: "seg0_stages_done", 12: "seg0_dfull", 13: "seg0_stored", 14: "dequant_done",
         16: "producer_start", 17: "tma_issued0", 18: "tma_issued1", 19: "tma_issued2", 48: "exit"}
for _i in range(4):
    NAMES[20 + 4 * _i] = f"st{_i}_x_staged"
    NAMES[21 + 4 * _i] = f"st{_i}_wfull"
    NAMES[22 + 4 * _i] = f"st{_i}_slot0"
    NAMES[23 + 4 * _i] = f"st{_i}_slot1"
NAMES[50] = "iss_before_dep"
NAMES[51] = "iss_dep_ok"
NAMES[52] = "iss_x_staged"
for _i in range(3):
    NAMES[36 + 3 * _i] = f"iss{_i}_xfull"
    NAMES[37 + 3 * _i] = f"iss{_i}_afull"
    NAMES[38 + 3 * _i] = f"iss{_i}_issued"


def main():
    lib = _native.capi()
    lib.tg_debug_set_trace.argtypes = [ctypes.c_void_p]
    lib.tg_set_option(1, int(os.environ.get("TR_STATIC", "1")))
    dev = torch.device("cuda:0")
    m = int(os.environ.get("TR_M", "1"))
    for n in [int(a) for a in sys.argv[1:]] or [4096, 8192]:
        k = n
        layers = [synth_layer(n, k, 10 + i, dev) for i in range(max(3, int(300e6 / (n * k / 2))))]
        x = torch.randn(m, k, device=dev).bfloat16()
        op = torch.ops.tinygemm.tinygemm_y_f16RM_x_f16RM_w_any4TC
        for w, lut, sz in layers:
            op(x, w, G, sz, lut, True)
        torch.cuda.synchronize()
        ctas = 512
        buf = torch.zeros(ctas * 64, dtype=torch.int64, device=dev)
        w, lut, sz = layers[0]
        w1, lut1, sz1 = layers[1]
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        buf0 = torch.zeros(ctas * 64, dtype=torch.int64, device=dev)
        if os.environ.get("TR_GRAPH", "0") == "1":
            # three GEMVs in one CUDA graph (programmatic edges): trace the 2nd and the 3rd
            w2, lut2, sz2 = layers[2]
            g = torch.cuda.CUDAGraph()
            with torch.cuda.graph(g):
                lib.tg_debug_set_trace(None)
                op(x, w2, G, sz2, lut2, True)
                lib.tg_debug_set_trace(ctypes.c_void_p(buf0.data_ptr()))
                op(x, w1, G, sz1, lut1, True)
                lib.tg_debug_set_trace(ctypes.c_void_p(buf.data_ptr()))
                op(x, w, G, sz, lut, True)
                lib.tg_debug_set_trace(None)
            e0.record()
            g.replay()
            e1.record()
            torch.cuda.synchronize()
        else:
            if os.environ.get("TR_PREV", "0") == "1":
                lib.tg_debug_set_trace(ctypes.c_void_p(buf0.data_ptr()))
                op(x, w1, G, sz1, lut1, True)
            lib.tg_debug_set_trace(ctypes.c_void_p(buf.data_ptr()))
            e0.record()
            op(x, w, G, sz, lut, True)
            e1.record()
            lib.tg_debug_set_trace(None)
            torch.cuda.synchronize()
        t = buf.cpu().numpy().reshape(ctas, 64).astype(np.int64)
        t = t[t[:, 0] > 0]
        ctas = len(t)
        t0 = t[:, 0].min()
        if os.environ.get("TR_PREV", "0") == "1" or os.environ.get("TR_GRAPH", "0") == "1":
            tp = buf0.cpu().numpy().reshape(-1, 64).astype(np.int64)
            tp = tp[tp[:, 0] > 0]
            print(f"previous kernel: entry {np.median(tp[:, 0] - t0) / 1e3:.2f} us, x staged {np.median(tp[:, 2] - t0) / 1e3:.2f}, "
                  f"exit median {np.median(tp[:, 48] - t0) / 1e3:.2f} max {(tp[:, 48].max() - t0) / 1e3:.2f} (relative to this kernel's first entry)")
        rel = (t - t0) / 1e3
        rel[t == 0] = np.nan
        print(f"== n=k={n} m={m}: {ctas} CTAs, event time {e0.elapsed_time(e1) * 1e3:.1f} us, SMs used {len(set(t[:, 63]))}")
        print(f"{'phase':18s} {'p10':>8s} {'median':>8s} {'p90':>8s} {'max':>8s}   (us since first CTA entry)")
        for i in sorted(NAMES):
            col = rel[:, i]
            if np.all(np.isnan(col)):
                continue
            q = np.nanpercentile(col, [10, 50, 90, 100])
            print(f"{NAMES[i]:18s} {q[0]:8.2f} {q[1]:8.2f} {q[2]:8.2f} {q[3]:8.2f}")
        dur = (t[:, 48] - t[:, 0]) / 1e3
        print(f"CTA lifetime: median {np.median(dur):.2f} us, min {dur.min():.2f}, max {dur.max():.2f}")


if __name__ == "__main__":
    main()
