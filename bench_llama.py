#!/usr/bin/env python
"""bench_llama.py - BASELINE.json configs[2] / configs[4]: Llama-3-8B any4 (g = 128, per-row LUT) single-token
decode, batch 1, synthetic weights / prompt, on 1 GPU or row-sharded over N GPUs (one NCCL all-reduce per Linear).

  python bench_llama.py [--steps K] [--warmup W] [--ctx L] [--layers 32]
  python -m torch.distributed.run --nproc-per-node N ... bench_llama.py --gpus N

The model is a bench harness around the hot path, not a product component: every nn.Linear of the decoder
stack (q, k, v, o, gate, up, down; `lm_head` stays bf16 as in the reference, quantize.py:34-36) is an
`any4_b200.modules.Any4Linear` filled with synthetic packed codes / LUTs / scales (the k-means quantizer is
out of scope and irrelevant for timing); attention (static KV cache of `--ctx` synthetic tokens, torch SDPA),
RMSNorm, RoPE and SiLU are stock torch ops.  One decode step is captured in a CUDA graph and replayed.
Architecture: public Llama-3-8B card (hidden 4096, 32 layers, 32 heads, 8 KV heads, intermediate 14336,
vocab 128256, rope theta 500000, rms eps 1e-5).
Reports tok/s = 1 / step latency and the fraction of the HBM roofline for the bytes a step must stream.
"""
import argparse
import json
import os
import sys

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

import torch  # noqa: E402
import torch.nn.functional as F  # noqa: E402

from bench import ClockSampler, G, measured_peak, synth_layer  # noqa: E402

HID, LAYERS, HEADS, KV_HEADS, INTER, VOCAB, HEAD_DIM = 4096, 32, 32, 8, 14336, 128256, 128
THETA, EPS = 500000.0, 1e-5


def any4_bytes(n, k):
    return n * k // 2 + (k // G) * n * 4 + n * 32


_FUSED = True


def make_linear(n, k, seed, dev, rank, world, fused=None):
    fused = _FUSED if fused is None else fused
    from any4_b200.modules import Any4Linear, RowShardedLinear

    lin = Any4Linear(k, n, bias=False, device=dev, dtype=torch.bfloat16, group_size=G)
    w, lut, sz = synth_layer(n, k, seed, dev)
    # keep activations tame through 32 random layers: centre the per-group zero, small scales
    lin.weight.data, lin.lut.data, lin.scales_and_zeros.data = w, lut, sz * 0.25
    lin.weight_reshaped = True
    return RowShardedLinear(lin, rank, world, fused=fused, max_features=INTER) if world > 1 else lin


class Block(torch.nn.Module):
    """Stock plumbing: separate q/k/v/gate/up launches, torch ops for norm / rope / attention / activation."""

    def __init__(self, idx, dev, rank, world, ctx):
        super().__init__()
        s = 1000 * idx
        self.q = make_linear(HID, HID, s + 1, dev, rank, world)
        self.k = make_linear(KV_HEADS * HEAD_DIM, HID, s + 2, dev, rank, world)
        self.v = make_linear(KV_HEADS * HEAD_DIM, HID, s + 3, dev, rank, world)
        self.o = make_linear(HID, HID, s + 4, dev, rank, world)
        self.gate = make_linear(INTER, HID, s + 5, dev, rank, world)
        self.up = make_linear(INTER, HID, s + 6, dev, rank, world)
        self.down = make_linear(HID, INTER, s + 7, dev, rank, world)
        self.n1 = torch.ones(HID, device=dev, dtype=torch.bfloat16)
        self.n2 = torch.ones(HID, device=dev, dtype=torch.bfloat16)
        gen = torch.Generator(device=dev).manual_seed(s + 9)
        self.kc = torch.randn(1, KV_HEADS, ctx + 1, HEAD_DIM, device=dev, generator=gen).bfloat16()
        self.vc = torch.randn(1, KV_HEADS, ctx + 1, HEAD_DIM, device=dev, generator=gen).bfloat16()
        self.ctx = ctx

    def forward(self, h, cos, sin):
        x = F.rms_norm(h, (HID,), self.n1, EPS)
        q = self.q(x).view(1, HEADS, 1, HEAD_DIM)
        k = self.k(x).view(1, KV_HEADS, 1, HEAD_DIM)
        v = self.v(x).view(1, KV_HEADS, 1, HEAD_DIM)

        def rope(t):
            t1, t2 = t[..., : HEAD_DIM // 2], t[..., HEAD_DIM // 2:]
            return t * cos + torch.cat((-t2, t1), -1) * sin

        q, k = rope(q), rope(k)
        self.kc[:, :, self.ctx:] = k
        self.vc[:, :, self.ctx:] = v
        a = F.scaled_dot_product_attention(q, self.kc, self.vc, enable_gqa=True)
        h = h + self.o(a.reshape(1, HID))
        x = F.rms_norm(h, (HID,), self.n2, EPS)
        return h + self.down(F.silu(self.gate(x)) * self.up(x))


class FusedBlock(torch.nn.Module):
    """any4_b200 plumbing: q|k|v and gate|up as one GEMV each (modules.fuse_rows) and three
    small PDL kernels (any4_b200.decode) - 8 launches per layer.  Same weights, same math as Block (the row-fused
    GEMVs are bit-identical to the separate ones; the element-wise kernels mirror the framework's roundings)."""

    def __init__(self, idx, dev, rank, world, ctx):
        super().__init__()
        from any4_b200.modules import RowShardedLinear, fuse_rows

        s = 1000 * idx
        mk = lambda n, k, seed: make_linear(n, k, seed, dev, rank, 1)  # noqa: E731  (shard after fusing)

        def shard(lin, cap):
            return RowShardedLinear(lin, rank, world, fused=_FUSED, max_features=cap) if world > 1 else lin

        self.qkv = shard(fuse_rows([mk(HID, HID, s + 1), mk(KV_HEADS * HEAD_DIM, HID, s + 2),
                                    mk(KV_HEADS * HEAD_DIM, HID, s + 3)]), 2 * INTER)
        self.o = shard(mk(HID, HID, s + 4), 2 * INTER)
        # 1 GPU: gate and up rows interleaved (synthetic packed data: any row order is the same workload) so that
        # silu(gate) * up happens in the GEMV epilogue;
        # sharded with the in-kernel exchange: the same, the shard keeps whole (gate, up) pairs; NCCL exchange: gate | up
        # blocks + the separate silu*mul kernel
        self.act_in_epilogue = world == 1 or _FUSED
        self.gate_up = (shard(mk(2 * INTER, HID, s + 5), 2 * INTER) if self.act_in_epilogue else
                        shard(fuse_rows([mk(INTER, HID, s + 5), mk(INTER, HID, s + 6)]), 2 * INTER))
        self.down = shard(mk(HID, INTER, s + 7), 2 * INTER)
        self.n1 = torch.ones(HID, device=dev, dtype=torch.bfloat16)
        self.n2 = torch.ones(HID, device=dev, dtype=torch.bfloat16)
        gen = torch.Generator(device=dev).manual_seed(s + 9)
        self.kc = torch.randn(1, KV_HEADS, ctx + 1, HEAD_DIM, device=dev, generator=gen).bfloat16()
        self.vc = torch.randn(1, KV_HEADS, ctx + 1, HEAD_DIM, device=dev, generator=gen).bfloat16()
        self.ctx = ctx

    def forward(self, h, delta, cos, sin):
        """h: residual stream (updated in place); delta: the previous layer's MLP output still to be added."""
        from any4_b200 import decode as D

        x = D.add_rmsnorm(h, delta, self.n1, EPS)
        a = D.rope_attention(self.qkv(x), cos, sin, self.kc, self.vc, self.ctx, HEADS, KV_HEADS, HEAD_DIM)
        x = D.add_rmsnorm(h, self.o(a), self.n2, EPS)
        act = D.linear_silu_pairs(self.gate_up, x) if self.act_in_epilogue else D.silu_mul(self.gate_up(x))
        return self.down(act)


class Llama(torch.nn.Module):
    def __init__(self, dev, rank, world, ctx, layers, fused_plumbing, lm_head_any4=False):
        super().__init__()
        gen = torch.Generator(device=dev).manual_seed(7)
        self.fused_plumbing = fused_plumbing
        self.emb = (torch.randn(VOCAB, HID, device=dev, generator=gen) * 0.02).bfloat16()
        blk = FusedBlock if fused_plumbing else Block
        self.blocks = torch.nn.ModuleList([blk(i, dev, rank, world, ctx) for i in range(layers)])
        self.norm = torch.ones(HID, device=dev, dtype=torch.bfloat16)
        # lm_head: bf16 on cuBLAS as in the reference (quantize.py:34-36 skips it), or any4 through the same GEMV kernel
        # (128256 rows, row-sharded like every other layer when world > 1)
        self.lm_head_any4 = lm_head_any4
        if lm_head_any4:
            self.lm_head = make_linear(VOCAB, HID, 777, dev, rank, world, fused=_FUSED)
            if world > 1:
                self.lm_head.max_features = max(self.lm_head.max_features, VOCAB)
        else:
            self.lm_head = (torch.randn(VOCAB, HID, device=dev, generator=gen) * 0.02).bfloat16()
        pos = torch.tensor([float(ctx)], device=dev)
        inv = 1.0 / (THETA ** (torch.arange(0, HEAD_DIM, 2, device=dev).float() / HEAD_DIM))
        ang = torch.cat([pos[:, None] * inv[None], pos[:, None] * inv[None]], -1)
        self.cos, self.sin = ang.cos().bfloat16().view(1, 1, 1, HEAD_DIM), ang.sin().bfloat16().view(1, 1, 1, HEAD_DIM)
        self.cos1, self.sin1 = self.cos.view(HEAD_DIM).contiguous(), self.sin.view(HEAD_DIM).contiguous()

    def forward(self, tok):
        h = self.emb[tok].view(1, HID)
        if self.fused_plumbing:
            from any4_b200 import decode as D

            delta = None
            for b in self.blocks:
                delta = b(h, delta, self.cos1, self.sin1)
            x = D.add_rmsnorm(h, delta, self.norm, EPS)
            return self.lm_head(x) if self.lm_head_any4 else F.linear(x, self.lm_head)
        for b in self.blocks:
            h = b(h, self.cos, self.sin)
        x = F.rms_norm(h, (HID,), self.norm, EPS)
        return self.lm_head(x) if self.lm_head_any4 else F.linear(x, self.lm_head)


def run_decode(dev, rank, world, dist, steps=50, warmup=5, ctx=128, layers=LAYERS, exchange="fused", plumbing="fused",
               lm_head_any4=False, reference=False, lib=None):
    """Build the synthetic model, capture one decode step in a CUDA graph, time `steps` replays (max over ranks).
    Returns the result dict on rank 0 (None elsewhere)."""
    local = dev.index

    def note(msg):
        if os.environ.get("LLAMA_DEBUG"):
            print(f"[rank {rank}] {msg}", file=sys.stderr, flush=True)

    if lib is not None:
        # TG_OPT_W4_KERNEL = 1: the tcgen05 kernel for every B-layout 4-bit GEMV.  The automatic choice sends the one
        # 4096 x 4096 projection to the mma.sync kernel, which wins in a chain of such GEMVs (bench.py's headline) but
        # needs a whole SM; between the small kernels of a decoder layer the 113 KB tcgen05 CTAs co-reside with their
        # neighbours and win (533 -> 555 tok/s on 1 GPU).
        lib.tg_set_option(2, 1)
    with torch.no_grad():
        note("building model")
        global _FUSED
        _FUSED = exchange == "fused"
        model = Llama(dev, rank, world, ctx, layers, plumbing == "fused", lm_head_any4)
        note("model built")
        tok = torch.tensor([1], device=dev)
        if lib is not None:
            lib.tg_reset_launch_count()
        logits = model(tok)
        torch.cuda.synchronize()
        note("first forward done")
        launches = int(lib.tg_launch_count()) if lib is not None else None
        assert torch.isfinite(logits.float()).all(), "synthetic model produced non-finite logits"
        for _ in range(warmup):
            model(tok)
        torch.cuda.synchronize()
        note("warm-up done, capturing")
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g):
            out = model(tok)
        note("captured")
        g.replay()
        torch.cuda.synchronize()
        note("first replay done")
        if dist is not None:
            dist.barrier()
        sampler = ClockSampler(local) if rank == 0 else None
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(steps):
            g.replay()
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / steps
        if dist is not None:
            t = torch.tensor([ms], device=dev, dtype=torch.float64)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms = float(t.item())
        clocks = sampler.stop() if sampler else None
        del g, out, model  # a live CUDA graph holding NCCL kernels keeps destroy_process_group() from returning
        if lib is not None:
            lib.tg_set_option(2, 0)
        torch.cuda.synchronize()
        torch.cuda.empty_cache()
    if rank != 0:
        return None
    per_layer = 2 * any4_bytes(HID, HID) + 2 * any4_bytes(KV_HEADS * HEAD_DIM, HID) + 2 * any4_bytes(INTER, HID) + any4_bytes(HID, INTER)
    quant_bytes = per_layer * layers
    head_bytes = any4_bytes(VOCAB, HID) / world if lm_head_any4 else VOCAB * HID * 2
    kv_bytes = layers * 2 * KV_HEADS * (ctx + 1) * HEAD_DIM * 2
    total = quant_bytes / world + head_bytes + kv_bytes  # per GPU: its shard of the any4 layers, replicated KV (and bf16 head)
    peak, src = measured_peak()
    return {
        "metric": "llama3_8b_any4_g128_decode_tok_per_s", "value": 1e3 / ms, "unit": "tok/s", "n_gpus": world,
        "impl": "reference tinygemm kernels (recompiled sm_100a) in the same harness" if reference else "any4_b200",
        "ms_per_token": ms, "steps": steps, "warmup": warmup, "dtype": "bf16", "data": "synthetic",
        "config": {"workload": "Llama-3-8B any4 g=128 single-token decode, batch 1 (BASELINE configs[2]/[4])",
                   "layers": layers, "kv_context": ctx, "launch": "one CUDA graph per token",
                   "parallelism": "1 GPU" if world == 1 else f"row-sharded x{world}, exchange per Linear: {exchange}",
                   "plumbing": ("q|k|v and gate|up row-fused GEMVs (silu*mul in the gate|up epilogue) + any4_b200.decode kernels, 7 launches / layer"
                                if plumbing == "fused" else "stock torch ops, 7 GEMV launches / layer"),
                   "lm_head": "any4 g=128 (row-sharded)" if lm_head_any4 else "bf16 (not quantized, as in the reference)",
                   "options": None if reference else "TG_OPT_STATIC_WEIGHTS = 1, TG_OPT_W4_KERNEL = 1 (tcgen05 kernel for every 4-bit GEMV)"},
        "bytes_per_token_per_gpu": total,
        "roofline": {"bound": "hbm", "achieved": total / (ms * 1e-3) / 1e9, "peak": peak, "unit": "GB/s",
                     "frac": total / (ms * 1e-3) / 1e9 / peak, "peak_source": src + " (of measured)"},
        "library_launches_per_token": launches, "clocks": clocks}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=50)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--ctx", type=int, default=128)
    ap.add_argument("--layers", type=int, default=LAYERS)
    ap.add_argument("--exchange", default="fused", choices=["fused", "nccl"])
    ap.add_argument("--plumbing", default="fused", choices=["fused", "torch"],
                    help="fused: q|k|v and gate|up as one GEMV each + any4_b200.decode kernels; torch: stock ops")
    ap.add_argument("--lm-head", default="bf16", choices=["bf16", "any4"],
                    help="bf16: cuBLAS as in the reference (which does not quantize lm_head); any4: through the GEMV kernel")
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"],
                    help="reference: the same harness (torch plumbing) on the UNMODIFIED reference tinygemm kernels "
                         "(oracle/_ref/tinygemm.so, recompiled for sm_100a); single GPU only")
    args = ap.parse_args()
    rank, world = int(os.environ.get("RANK", "0")), int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist = None
    if world > 1:
        import torch.distributed as dist_mod

        dist_mod.init_process_group("nccl", device_id=dev)
        dist = dist_mod
    reference = args.impl == "reference"
    if reference:
        assert world == 1, "--impl reference is single-GPU"
        ref_so = os.path.join(ROOT, "oracle", "_ref", "tinygemm.so")
        if not os.path.exists(ref_so):
            print(json.dumps({"impl": "reference", "unavailable": "oracle/_ref/tinygemm.so not built"}))
            return
        # benchmarking aid: register the unmodified reference extension's `tinygemm::` ops instead of this repo's and
        # tell the package loader they are there; any4_b200.modules / functional (pure Python) then drive them
        torch.ops.load_library(ref_so)
        from any4_b200 import _native as _nat
        _nat._ops_loaded = True
        args.plumbing = "torch"
    from any4_b200 import _native
    from any4_b200 import functional as tgf

    lib = None
    if not reference:
        tgf.set_static_weights(True)
        lib = _native.capi()
    res = run_decode(dev, rank, world, dist, args.steps, args.warmup, args.ctx, args.layers, args.exchange, args.plumbing,
                     args.lm_head == "any4", reference, lib)
    if rank == 0:
        print(json.dumps(res))
    if dist is not None:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
