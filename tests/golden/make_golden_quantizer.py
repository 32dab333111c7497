#!/usr/bin/env python
"""Generate tests/golden/golden_quantizer.npz by IMPORTING THE REFERENCE (build container only): the any4 quantizer
front-end on seeded weights.
  * quantize.group_q                                  -> v (scaled values), scales_and_zeros
  * kmeans.run_kmeans(init = build_init(.., "int"))    -> the reference's own Lloyd on each row (labels, centroids)
  * quantize.anyq_quantize_tensor(.., init="int")      -> the default sklearn path: the quality bar (reconstruction MSE)
bf16 tensors are stored as uint16 bit patterns (key suffix `__bf16`)."""
import os
import sys
import types

import numpy as np
import torch

REF = os.environ.get("ANY4_REFERENCE", "/root/reference")
sys.path.insert(0, REF)
sys.modules.setdefault("bitsandbytes", types.ModuleType("bitsandbytes"))

import kmeans as refk  # noqa: E402
import quantize as refq  # noqa: E402

out = {}


def put(name, t):
    if isinstance(t, torch.Tensor):
        if t.dtype == torch.bfloat16:
            out[name + "__bf16"] = t.contiguous().view(torch.int16).numpy().view(np.uint16)
            return
        t = t.contiguous().numpy()
    out[name] = np.asarray(t)


torch.manual_seed(4321)
np.random.seed(0)
case = 0
for (n, k, g, weighted) in [(16, 512, 128, False), (8, 1024, 64, True), (8, 256, 32, False), (4, 2048, 256, False)]:
    W = (torch.randn(n, k) * 0.05 * (1 + 3 * torch.rand(n, 1))).bfloat16()
    W[0, :7] *= 8  # a few outliers
    sw = (torch.rand(k) + 0.1) if weighted else None
    v, _, sz = refq.group_q(W, 4, q_group_size=g)
    put(f"q{case}_w", W)
    put(f"q{case}_v", v)
    put(f"q{case}_sz", sz.to(W.dtype))
    if sw is not None:
        put(f"q{case}_sw", sw)
    labels = np.zeros((n, k), dtype=np.int32)
    cen = np.zeros((n, 16), dtype=np.float32)
    vn = v.float().numpy()
    for r in range(n):
        x = vn[r].reshape(-1, 1)
        init = refk.build_init(x=x, n_clusters=16, init_type="int").numpy()
        weight = None if sw is None else refk.build_sample_weight(x=x, sample_weight_type=sw.numpy(), abs=True)
        weight = np.ones(x.shape[0]) if weight is None else weight   # as kmeans.kmeans does (kmeans.py:168-169)
        # (kmeans.kmeans itself cannot take an array init: `init in [...]` / `init == 'k-means++'` are ambiguous for
        # arrays, kmeans.py:172, :213 - so its Lloyd loop is called directly, with the tolerance kmeans() would pass)
        _, c, lab = refk.run_kmeans(x, init.astype(np.float32).copy(), 300, refk._tolerance(x, 1e-4), 0, weight)
        labels[r], cen[r] = lab, c.reshape(16)
    put(f"q{case}_labels", labels)
    put(f"q{case}_centroids", cen)
    assign, any4, sz2 = refq.anyq_quantize_tensor(W, n_bit=4, q_group_size=g, per_row=True, init="int", sample_weight=sw,
                                                  parallelize=False)
    Wd = refq.anyq_dequantize_tensor(assign, any4, sz2, n_bit=4, q_group_size=g, per_row=True)
    out[f"q{case}_sklearn_mse"] = np.array(float(((Wd.float() - W.float()) ** 2).mean()))
    out[f"q{case}_meta"] = np.array([n, k, g, int(weighted)])
    case += 1
out["n_cases"] = np.array(case)
path = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden_quantizer.npz")
np.savez_compressed(path, **out)
print("wrote", path, {k: v.shape for k, v in out.items() if k.endswith("labels")}, [float(out[f"q{i}_sklearn_mse"]) for i in range(case)])
