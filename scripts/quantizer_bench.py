#!/usr/bin/env python
"""Time of the GPU any4 quantizer front-end (one kernel per weight matrix) on the Llama-3-8B layer shapes."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from any4_b200.quantize import anyq_quantize_tensor  # noqa: E402

dev = torch.device("cuda:0")
total = 0.0
for name, n, k, count in (("q/o 4096x4096", 4096, 4096, 2), ("k/v 1024x4096", 1024, 4096, 2), ("gate/up 14336x4096", 14336, 4096, 2),
                          ("down 4096x14336", 4096, 14336, 1), ("lm_head 128256x4096", 128256, 4096, 0)):
    W = (torch.randn(n, k, device=dev) * 0.02).bfloat16()
    for _ in range(2):
        out = anyq_quantize_tensor(W, pack_inner_k=4)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    out = anyq_quantize_tensor(W, pack_inner_k=4)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1)
    total += ms * count
    print(f"{name}: {ms:.2f} ms")
    del W, out
print(f"one decoder layer: {total:.1f} ms; 32 layers: {32 * total / 1e3:.2f} s")
