#!/usr/bin/env python
"""us per GEMM of the int8 / 16-bit weight kernels (B layout, m = KB_M, n = k = 4096) - tuning / ncu helper."""
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import tinygemm  # noqa: E402,F401
from bench import G, synth_layer  # noqa: E402


def main():
    dev = torch.device("cuda:0")
    ops = torch.ops.tinygemm
    n = k = int(os.environ.get("KB_N", "4096"))
    m = int(os.environ.get("KB_M", "1"))
    gen = torch.Generator(device=dev).manual_seed(1)
    x = torch.randn(m, k, device=dev).bfloat16()
    sz = synth_layer(n, k, 1, dev)[2]
    w8 = [torch.randint(-2**31, 2**31 - 1, (n // 8, k // 64, 32, 4), generator=gen, device=dev, dtype=torch.int64).to(torch.int32)
          for _ in range(16)]
    w16 = [torch.randn(n // 8, k // 32, 32, 8, generator=gen, device=dev).bfloat16() for _ in range(8)]
    out = {}
    for name, fn, copies in (("int8", lambda: [ops.tinygemm_y_f16RM_x_f16RM_w_int8TC(x, w, G, sz, True) for w in w8], 16),
                             ("bf16", lambda: [ops.tinygemm_y_f16RM_x_f16RM_w_f16TC(x, w, True) for w in w16], 8)):
        for _ in range(3):
            fn()
        torch.cuda.synchronize()
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g):
            fn()
        g.replay()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(10):
            g.replay()
        e1.record()
        torch.cuda.synchronize()
        out[name] = round(e0.elapsed_time(e1) * 1e3 / (10 * copies), 2)
    print(json.dumps({"m": m, "n": n, "us_per_gemm": out}))


if __name__ == "__main__":
    main()
