timeout 900 python -m pytest tests/test_gemm_gpu.py tests/test_vs_reference_gpu.py tests/test_modules_gpu.py tests/test_capi_gpu.py -x -q 2>&1 | tail -4
timeout 200 python - <<'PY'
import torch, tinygemm
ops = torch.ops.tinygemm
dev = torch.device("cuda:0"); n = k = 4096; G = 128
gen = torch.Generator(device=dev).manual_seed(1)
def timed(fn, copies):
    for _ in range(2): fn()
    torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g): fn()
    g.replay(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(5): g.replay()
    e1.record(); torch.cuda.synchronize()
    return round(e0.elapsed_time(e1) * 1e3 / (5 * copies), 2)
sz = torch.stack([torch.rand(k // G, n, generator=gen, device=dev) * 0.01 + 0.001, torch.randn(k // G, n, generator=gen, device=dev) * 0.01], 2).bfloat16().contiguous()
out = {}
for m in (1, 4, 16):
    x = torch.randn(m, k, device=dev).bfloat16()
    w8b = [torch.randint(-2**31, 2**31 - 1, (n // 8, k // 64, 32, 4), generator=gen, device=dev, dtype=torch.int64).to(torch.int32) for _ in range(16)]
    out[f"int8_B_m{m}"] = timed(lambda: [ops.tinygemm_y_f16RM_x_f16RM_w_int8TC(x, w, G, sz, True) for w in w8b], 16)
    w8a = [w.view(n // 16, k // 32, 32, 4) for w in w8b]
    out[f"int8_A_ik2_m{m}"] = timed(lambda: [ops.tinygemm_y_f16RM_x_f16RM_w_int8TC(w, x, G, sz, False) for w in w8a], 16)
    del w8b, w8a
    w16 = [torch.randn(n // 8, k // 32, 32, 8, generator=gen, device=dev).bfloat16() for _ in range(8)]
    out[f"bf16_B_m{m}"] = timed(lambda: [ops.tinygemm_y_f16RM_x_f16RM_w_f16TC(x, w, True) for w in w16], 8)
    w16a = [w.view(n // 16, k // 16, 32, 8) for w in w16]
    out[f"bf16_A_m{m}"] = timed(lambda: [ops.tinygemm_y_f16RM_x_f16RM_w_f16TC(w, x, False) for w in w16a], 8)
    del w16, w16a
print(out)
PY
