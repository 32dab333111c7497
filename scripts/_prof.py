import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, tinygemm
ops = torch.ops.tinygemm
dev = torch.device("cuda:0"); n = k = 4096; G = 128
gen = torch.Generator(device=dev).manual_seed(1)
sz = torch.stack([torch.rand(k // G, n, generator=gen, device=dev) * 0.01 + 0.001, torch.randn(k // G, n, generator=gen, device=dev) * 0.01], 2).bfloat16().contiguous()
x = torch.randn(1, k, device=dev).bfloat16()
w8b = [torch.randint(-2**31, 2**31 - 1, (n // 8, k // 64, 32, 4), generator=gen, device=dev, dtype=torch.int64).to(torch.int32) for _ in range(16)]
for _ in range(2):
    for w in w8b:
        ops.tinygemm_y_f16RM_x_f16RM_w_int8TC(x, w, G, sz, True)
torch.cuda.synchronize()
