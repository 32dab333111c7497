"""TEST INFRASTRUCTURE - CPU restatement (numpy, float32) of the reference's any4 quantizer front-end for the
configuration the any4 Linear uses: 4 bit, per-row LUT, asymmetric groups with zero point, "int" initialisation.

  group_q          quantize.py:106-149    scale = clamp(max - min, 1e-6) / 15, zero = min + 8 * scale, v = (w - min) / scale
  build_init "int" kmeans.py:41-46        torch.linspace(min(v), max(v), 16) per row
  run_kmeans       kmeans.py:200-262      Lloyd: nearest centroid (argmin, ties -> lower index); stop when the labels repeat
                                          (before updating), weighted mean per cluster (plain mean when the weights sum
                                          to 0, empty clusters keep their centroid), stop when i > 0 and
                                          ||c - c_old|| < mean(var(v)) * tol (kmeans.py:186-196), at most max_iter rounds
  lut              quantize.py:893        any4.to(dtype) - 8 in dtype
Pinned by tests/golden/golden_quantizer.npz (tests/golden/make_golden_quantizer.py imports the real reference:
kmeans.kmeans with the same init, and the default sklearn path for the quality bound).  Only tests may import this.
"""
import numpy as np
import torch


def group_q(w, group):
    """w [n][k] torch (any float dtype) -> v [n][k] float32 in [0, 15], scales_and_zeros [k/g][n][2] in w.dtype"""
    n, k = w.shape
    x = w.float().reshape(-1, group)
    mx, mn = x.amax(1, keepdim=True), x.amin(1, keepdim=True)
    scale = (mx - mn).clamp(min=1e-6) / 15
    zero = mn + scale * 8
    v = x.sub(mn).div(scale).reshape(n, k)
    sz = torch.cat([scale.reshape(n, -1, 1), zero.reshape(n, -1, 1)], 2).transpose(0, 1).contiguous()  # pack_scales_and_zeros
    return v, sz.to(w.dtype)


def linspace16(lo, hi):
    """torch.linspace(lo, hi, 16) in float32"""
    return torch.linspace(float(lo), float(hi), 16, dtype=torch.float32).numpy()


def lloyd_row(x, sample_weight=None, max_iter=300, tol=1e-4):
    """x [k] float32 -> (labels [k] int, centroids [16] float32, iterations)"""
    x = np.ascontiguousarray(x, dtype=np.float32)
    sw = np.ones_like(x) if sample_weight is None else np.asarray(sample_weight, dtype=np.float32)
    cen = linspace16(x.min(), x.max()).astype(np.float32)
    thr = np.float32(np.var(x) * tol)
    labels = np.zeros(x.shape[0], dtype=np.int64)
    old = None
    it = 0
    for it in range(max_iter):
        new = np.argmin(np.abs(x[:, None] - cen[None, :]), axis=1)
        if np.array_equal(new, labels):
            break
        labels = new
        for j in range(16):
            m = labels == j
            if m.any():
                w = sw[m]
                cen[j] = np.float32(np.average(x[m]) if w.sum() == 0 else np.average(x[m], weights=w))
        if it > 0 and np.linalg.norm(cen - old) < thr:
            it += 1
            break
        old = cen.copy()
    else:
        it = max_iter
    return labels, cen, it


def quantize_any4(w, group=128, sample_weight=None, max_iter=300, tol=1e-4):
    """w [n][k] torch bf16/fp16 -> dict(codes [n][k] int32, any4 [n][16] dtype, lut [n][16] dtype, sz [k/g][n][2] dtype)"""
    v, sz = group_q(w, group)
    n, k = w.shape
    codes = np.zeros((n, k), dtype=np.int32)
    cen = np.zeros((n, 16), dtype=np.float32)
    iters = np.zeros(n, dtype=np.int32)
    sw = None if sample_weight is None else sample_weight.float().numpy()
    vn = v.numpy()
    for r in range(n):
        codes[r], cen[r], iters[r] = lloyd_row(vn[r], sw, max_iter, tol)
    any4 = torch.from_numpy(cen).to(w.dtype)
    lut = any4 - 8
    return dict(codes=torch.from_numpy(codes), any4=any4, lut=lut, sz=sz, v=v, iters=torch.from_numpy(iters))


def dequantize(codes, any4, sz, group):
    """quantize.py:612-637 (centering applied to the un-centred table): W = (any4[row][code] - 8) * scale + zero, fp32"""
    n, k = codes.shape
    scale = sz[..., 0].t().float().repeat_interleave(group, 1)
    zero = sz[..., 1].t().float().repeat_interleave(group, 1)
    val = torch.gather(any4.float(), 1, codes.long())
    return (val - 8) * scale + zero
