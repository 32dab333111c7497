import sys, os, torch
sys.path.insert(0, ".")
import tinygemm
from any4_b200 import _native
from bench import synth_layer
lib = _native.capi()
dev = torch.device("cuda:0")
ops = torch.ops.tinygemm
n = k = 4096
gen = torch.Generator(device=dev).manual_seed(1)
ws = [synth_layer(n, k, 500 + i, dev)[0] for i in range(37)]
exps = torch.randint(118, 130, (n, k // 32), generator=gen, device=dev, dtype=torch.int32).to(torch.uint8)
sz32 = torch.stack([torch.rand(k // 32, n, generator=gen, device=dev) * 0.01 + 0.001, torch.randn(k // 32, n, generator=gen, device=dev) * 0.01], 2).bfloat16().contiguous()
lut = (torch.rand(n, 16, device=dev, generator=gen) * 15).sort(1).values.bfloat16() - 8
from any4_b200 import functional as tgf
tgf.set_static_weights(True)
for m in (1,):
    x = torch.randn(m, k, device=dev).bfloat16()
    for name, fn in (("mx4 g32", lambda: [ops.tinygemm_y_f16RM_x_f16RM_w_mx4TC(x, w, 32, exps, True) for w in ws]),
                     ("any4 g32", lambda: [ops.tinygemm_y_f16RM_x_f16RM_w_any4TC(x, w, 32, sz32, lut, True) for w in ws])):
        for kern in (1, 2):
            lib.tg_set_option(2, kern)
            for _ in range(2): fn()
            torch.cuda.synchronize()
            g = torch.cuda.CUDAGraph()
            with torch.cuda.graph(g): fn()
            g.replay(); torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(5): g.replay()
            e1.record(); torch.cuda.synchronize()
            print(name, "m", m, "kernel", kern, round(e0.elapsed_time(e1) * 1e3 / (5 * 37), 2), "us")
lib.tg_set_option(2, 0)
