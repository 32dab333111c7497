#!/usr/bin/env python
"""us per GEMM of the A-layout int4 op (Int4Linear's default kernel) at 4096^2 for several activation row counts."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import tinygemm  # noqa: E402,F401
from bench import G, synth_layer  # noqa: E402

dev = torch.device("cuda:0")
ops = torch.ops.tinygemm
n = k = 4096
ws = [synth_layer(n, k, 500 + i, dev) for i in range(37)]
wa = [w.view(n // 16, k // 64, 32, 4) for w, _, _ in ws]
for m in (1, 2, 3, 4, 8, 16):
    x = torch.randn(m, k, device=dev).bfloat16()
    fn = lambda: [ops.tinygemm_y_f16RM_x_f16RM_w_int4TC(w, x, G, sz, False) for w, (_, _, sz) in zip(wa, ws)]
    for _ in range(2):
        fn()
    torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        fn()
    g.replay()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(5):
        g.replay()
    e1.record()
    torch.cuda.synchronize()
    print("A layout m =", m, round(e0.elapsed_time(e1) * 1e3 / (5 * 37), 2), "us")
