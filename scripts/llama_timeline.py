#!/usr/bin/env python
"""Kernel timeline of one CUDA-graph decode step of bench_llama.py's model (torch profiler / CUPTI): start offset,
duration and the gap to the previous kernel's end, for the first `--show` kernels after the first layer.  Tuning aid."""
import argparse
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--layers", type=int, default=4)
    ap.add_argument("--plumbing", default="fused")
    ap.add_argument("--show", type=int, default=40)
    args = ap.parse_args()
    import bench_llama as B
    from any4_b200 import functional as tgf

    tgf.set_static_weights(True)
    dev = torch.device("cuda:0")
    with torch.no_grad():
        model = B.Llama(dev, 0, 1, 128, args.layers, args.plumbing == "fused")
        tok = torch.tensor([1], device=dev)
        for _ in range(3):
            model(tok)
        torch.cuda.synchronize()
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g):
            model(tok)
        for _ in range(3):
            g.replay()
        torch.cuda.synchronize()
        from torch.profiler import ProfilerActivity, profile

        with profile(activities=[ProfilerActivity.CUDA, ProfilerActivity.CPU]) as prof:
            g.replay()
            torch.cuda.synchronize()
    evs = [e for e in prof.events() if e.device_type == torch.autograd.DeviceType.CUDA and "emcpy" not in e.name]
    evs.sort(key=lambda e: e.time_range.start)
    t0 = evs[0].time_range.start
    prev_end = t0
    print(f"{len(evs)} kernels, step span {(evs[-1].time_range.end - t0):.1f} us")
    for e in evs[: args.show]:
        s, d = e.time_range.start - t0, e.time_range.end - e.time_range.start
        print(f"{s:9.2f} us  dur {d:7.2f}  gap {e.time_range.start - prev_end:7.2f}  {e.name[:90]}")
        prev_end = max(prev_end, e.time_range.end)


if __name__ == "__main__":
    main()
