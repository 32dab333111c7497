#!/usr/bin/env python
"""bench.py - the hot-path benchmark (BASELINE.json configs[1]):
tinygemm any4-bf16 GEMV, m = 1, n = k = 4096 (headline) plus 8192 and 11008, g = 128.

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]

One "step" = one GEMV on each of `copies` distinct synthetic weight sets of the headline shape
(the copies total > 2.5x the 126 MB L2, so every launch streams its weights from HBM).
value   = algorithmic GB/s (SURVEY.md 8d byte count) with everything resident in HBM
e2e     = same metric through the public module API (any4_b200.modules.Any4Linear) with HOST
          buffers: pinned x -> device, GEMV, y -> host, every call
N > 1   = the weight rows are sharded across ranks (RowShardedLinear), one NCCL all-reduce on
          the m x n output per GEMV; value = total algorithmic bytes / max-over-ranks time
--impl reference times the reference's own CPU path (quantize.py:612-637 dequant + F.linear,
restated in oracle/cpu_path.py) on the host cores for the same config.
Prints ONE JSON line on rank 0.
"""
import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

import torch  # noqa: E402

G = 128
HEADLINE = 4096
EXTRA_SHAPES = (8192, 11008)
METRIC = "any4_gemv_m1_n4096_k4096_g128_algorithmic_GBps"
WORKLOAD = "any4-bf16 GEMV m=1 n=k=4096 g=128 per-row LUT (BASELINE configs[1])"  # the same string in both arms


def algorithmic_bytes(n, k, m=1, g=G):
    """packed weights + scale/zero + per-row LUT + x + y (bf16)"""
    return n * k // 2 + (k // g) * n * 4 + n * 32 + 2 * m * k + 2 * m * n


def measured_peak():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            return float(json.load(open(p))["hbm_gbs"]), "MEASURED_PEAKS.json hbm_gbs"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md)"


def ncu_traffic():
    """DRAM bytes per launch of the headline kernel (dram__bytes_read.sum + dram__bytes_write.sum) from the committed
    `ncu --set full` capture of this same kernel and shape taken in this round's GPU run (profiles/r2/ncu_b4096.txt, stamped
    with the commit it was taken at); None if absent.  (ncu cannot run inside the timed bench: a profiled run is never timed.)"""
    path = os.path.join(ROOT, "profiles", "r2", "ncu_b4096.txt")
    scale = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
    try:
        tot = 0.0
        for line in open(path):
            f = line.rstrip("\n").split("\t")
            if f[0] in ("dram__bytes_read.sum", "dram__bytes_write.sum"):
                vals = [float(v) for v in f[2:] if v]
                tot += scale[f[1]] * sum(vals) / len(vals)
        return tot or None
    except Exception:
        return None


def synth_layer(n, k, seed, device, rows=None):
    """Synthetic any4 layer (SURVEY.md 8d recipe), generated on the device: packed codes are
    uniformly random nibbles, so the packed words are drawn directly (equivalent to packing
    randint(0,16) codes, and 8x less memory than the int32 code matrix)."""
    gen = torch.Generator(device=device).manual_seed(seed)
    rows = n if rows is None else rows
    w = torch.randint(-2**31, 2**31 - 1, (rows // 8, k // 64, 32, 2), generator=gen, device=device,
                      dtype=torch.int64).to(torch.int32)
    lut = ((torch.rand(rows, 16, generator=gen, device=device) * 15).sort(1).values.bfloat16() - 8)
    scale = torch.rand(k // G, rows, generator=gen, device=device) * 0.01 + 0.001
    zero = torch.randn(k // G, rows, generator=gen, device=device) * 0.01
    sz = torch.stack([scale, zero], dim=2).bfloat16().contiguous()
    return w, lut, sz


class ClockSampler:
    """nvidia-smi clocks / throttle reasons during the timed region (B200_PROFILING.md recipe)"""

    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index=0):
        self.rows = []
        self.proc = None
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", "-i", str(index), f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "100"],
                stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append(line.strip())

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in self.rows:
            f = [v.strip() for v in r.split(",")]
            try:
                sm.append(float(f[0])); mx.append(float(f[1]))
            except Exception:
                continue
            for nme, v in zip(names, f[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(nme)
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "samples": len(sm), "reasons": sorted(reasons)}


def time_gemv_set(call, n_calls_per_step, steps, warmup, dist=None, graph=False):
    """W untimed steps, then exactly K timed steps bracketed by barrier + synchronize; CUDA events
    on the launching (current) stream; returns max-over-ranks milliseconds for the K steps.
    graph=True captures one step in a CUDA graph (after the eager warm-up) and replays it K times:
    the ~2 us GEMV is shorter than an eager launch, so only a graph measures the GPU, not the host."""
    for _ in range(warmup):
        call()
    torch.cuda.synchronize()
    if graph:
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g):
            call()
        eager_call, call = call, g.replay
        call()
        torch.cuda.synchronize()
    if dist is not None:
        dist.barrier()
        torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(steps):
        call()
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1)
    if dist is not None:
        t = torch.tensor([ms], device="cuda", dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        dist.barrier()
        ms = float(t.item())
    return ms


def cpu_baseline(seconds=12.0):
    """The reference's CPU path (dense dequant + F.linear) on the host cores, bounded sample."""
    from oracle import cpu_path

    try:
        torch.set_num_threads(max(1, len(os.sched_getaffinity(0))))
    except Exception:
        pass

    n = k = HEADLINE
    gen = torch.Generator().manual_seed(0)
    assign = torch.randint(0, 16, (n, k), generator=gen, dtype=torch.int32)
    any4 = (torch.rand(n, 16, generator=gen) * 15).sort(1).values.bfloat16()
    sz = torch.stack([torch.rand(k // G, n, generator=gen) * 0.01 + 0.001,
                      torch.randn(k // G, n, generator=gen) * 0.01], dim=2).bfloat16()
    x = torch.randn(1, k, generator=gen).bfloat16()
    for _ in range(2):
        cpu_path.any4_linear_forward(x, assign, any4, sz, G, True)
    t0 = time.perf_counter()
    reps = 0
    while True:
        cpu_path.any4_linear_forward(x, assign, any4, sz, G, True)
        reps += 1
        if time.perf_counter() - t0 > seconds or reps >= 200:
            break
    dt = (time.perf_counter() - t0) / reps
    return {"value": algorithmic_bytes(n, k) / dt / 1e9, "unit": "GB/s", "cores": torch.get_num_threads(),
            "kind": "port", "ms_per_forward": dt * 1e3,
            "sample": f"{reps} forwards of one 4096x4096 g=128 m=1 any4 Linear (dense dequant + F.linear, bf16)"}


def run_reference(args):
    """--impl reference: the reference's CPU implementation of the path, all host threads."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    from oracle import cpu_path

    # all the host threads the box offers - torchrun exports OMP_NUM_THREADS=1 to its workers, which would starve this arm
    try:
        torch.set_num_threads(max(1, len(os.sched_getaffinity(0))))
    except Exception:
        pass
    n = k = HEADLINE
    per_step = 2  # bounded sample: 2 forwards per step (ours: `copies` GEMVs per step)
    gen = torch.Generator().manual_seed(0)
    assign = torch.randint(0, 16, (n, k), generator=gen, dtype=torch.int32)
    any4 = (torch.rand(n, 16, generator=gen) * 15).sort(1).values.bfloat16()
    sz = torch.stack([torch.rand(k // G, n, generator=gen) * 0.01 + 0.001,
                      torch.randn(k // G, n, generator=gen) * 0.01], dim=2).bfloat16()
    x = torch.randn(1, k, generator=gen).bfloat16()

    def step():
        for _ in range(per_step):
            cpu_path.any4_linear_forward(x, assign, any4, sz, G, True)

    for _ in range(args.warmup):
        step()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        step()
    dt = time.perf_counter() - t0
    value = algorithmic_bytes(n, k) * per_step * args.steps / dt / 1e9
    cores = torch.get_num_threads()
    sample = f"{per_step} forwards/step of the 4096x4096 g=128 m=1 any4 Linear on CPU (dense dequant + F.linear)"
    print(json.dumps({
        "impl": "reference", "metric": METRIC, "value": value, "unit": "GB/s", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": dt * 1e3 / args.steps, "higher_is_better": True,
        "scaling": "strong", "vs_baseline": None, "dtype": "bf16", "data": "synthetic",
        "config": {"workload": WORKLOAD, "device": "cpu"},
        "cpu_baseline": {"value": value, "unit": "GB/s", "cores": cores, "kind": "port", "sample": sample},
        "e2e": {"value": value, "unit": "GB/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }))


def run_ours(args):
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device - the product path has no CPU fallback")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist = None
    if world > 1:
        import torch.distributed as dist_mod

        dist_mod.init_process_group("nccl", device_id=dev)
        dist = dist_mod

    from any4_b200 import _native
    from any4_b200.modules import Any4Linear, RowShardedLinear

    lib = _native.capi()
    peak, peak_src = measured_peak()
    from any4_b200 import functional as tgf

    tgf.set_static_weights(True)  # every layer below is packed once, before the timed region

    parity = {"checked": world > 1, "ok": True}

    def build_layers(n, k, copies):
        layers = []
        for i in range(copies):
            lin = Any4Linear(k, n, bias=False, device=dev, dtype=torch.bfloat16, group_size=G)
            w, lut, sz = synth_layer(n, k, 1234 + i, dev)
            lin.weight.data, lin.lut.data, lin.scales_and_zeros.data = w, lut, sz
            lin.weight_reshaped = True
            if world > 1:
                sh = RowShardedLinear(lin, rank, world, fused=args.exchange == "fused", max_features=11008)
                if i == 0:
                    # driver-visible parity of the N > 1 path, before anything is timed: the sharded output (every
                    # rank's copy) against the single-GPU output of the same layer, m = 1 and m = 5.  Same dequantised
                    # weights and exact products; a shard of n / N rows may get another k-split than the full layer, so
                    # the fp32 summation ORDER (the last bf16 bit of a few outputs) may differ - nothing else may.
                    for m in (1, 5):
                        xp = torch.randn(m, k, device=dev, generator=torch.Generator(device=dev).manual_seed(5 + m)).bfloat16()
                        got, want = sh(xp).float(), lin(xp).float()   # (every rank makes the same calls, whatever `ok` is)
                        close = bool(((got - want).abs() <= 2.0 ** -7 * want.abs() + 2.0 ** -9 * want.abs().max()).all())
                        same = float((got == want).float().mean())
                        parity["ok"] = parity["ok"] and close and same > 0.98
                        parity["frac_bit_equal"] = min(parity.get("frac_bit_equal", 1.0), same)
                layers.append(sh)
            else:
                layers.append(lin)
        return layers

    def bench_shape(n, k, steps, warmup, with_e2e):
        nbytes = algorithmic_bytes(n, k)
        copies = max(3, int(2.6 * 126e6 / nbytes) + 1)
        layers = build_layers(n, k, copies)
        x = torch.randn(1, k, device=dev).bfloat16()
        ys = [None]

        def step():
            for lin in layers:
                ys[0] = lin(x)

        lib.tg_reset_launch_count()
        step()
        launches = int(lib.tg_launch_count()) * steps  # our kernels per step (same under graph replay) x K
        ms = time_gemv_set(step, copies, steps, warmup, dist, graph=not args.no_graph)
        out = {"n": n, "k": k, "copies": copies, "ms": ms, "launches": launches,
               "us_per_gemv": ms * 1e3 / (steps * copies),
               "gbps": nbytes * copies * steps / (ms * 1e-3) / 1e9}
        if with_e2e and world == 1 and not args.no_graph:
            # informational: the SAME GEMVs issued as two independent stream-ordered chains (even / odd weight sets on two
            # streams, forked and joined inside one CUDA graph).  `value` above is ONE dependent chain, where every GEMV
            # waits for the previous one (each launch pays the dependency-resolution and prologue latency in full); two
            # chains let the hardware overlap one GEMV's latency with the other's dequant.  Not the headline.
            s2 = [torch.cuda.Stream(), torch.cuda.Stream()]
            outs2 = [None] * len(layers)

            def step2():
                cur = torch.cuda.current_stream()
                for j, st in enumerate(s2):
                    st.wait_stream(cur)
                    with torch.cuda.stream(st):
                        for i in range(j, len(layers), 2):
                            outs2[i] = layers[i](x)
                for st in s2:
                    cur.wait_stream(st)

            ms3 = time_gemv_set(step2, copies, steps, warmup, dist, graph=True)
            out["two_chains_us_per_gemv"] = ms3 * 1e3 / (steps * copies)
            out["two_chains_gbps"] = nbytes * copies * steps / (ms3 * 1e-3) / 1e9
        if with_e2e:
            # end to end through the public module API with HOST buffers: the activations start in pinned host memory
            # and the outputs must be readable on the host when the step ends (synchronize inside the timed region)
            xh = torch.randn(1, k).bfloat16().pin_memory()
            if world == 1:
                # Any4Linear.bind_host: per call a one-CTA kernel pulls x (k * 2 bytes) out of pinned host memory and the
                # GEMV kernel writes y (n * 2 bytes) straight into the pinned host buffer - no copy-engine transfers
                bound = [lin.bind_host(xh) for lin in layers]
                out["e2e_api"] = "Any4Linear.bind_host (tg_gemm_w4_rm_hostio: x pulled from / y written to pinned host memory by the kernels)"

                def enqueue_e2e():
                    for launch, _ in bound:
                        launch()

                def step_e2e_eager():
                    enqueue_e2e()
                    torch.cuda.synchronize()

                # the step's launches captured once in a CUDA graph, like the device-resident measurement above: per
                # step ONE graph launch; the copies over PCIe and the host synchronisation stay inside the timed region
                for _ in range(2):
                    step_e2e_eager()
                g_e2e = torch.cuda.CUDAGraph()
                with torch.cuda.graph(g_e2e):
                    enqueue_e2e()

                def step_e2e():
                    g_e2e.replay()
                    torch.cuda.synchronize()

                ms_eager = time_gemv_set(step_e2e_eager, copies, steps, warmup, dist)
                out["e2e_eager_gbps"] = nbytes * copies * steps / (ms_eager * 1e-3) / 1e9
                out["e2e_api"] += "; the step's launches replayed as one CUDA graph"
            else:
                yh = torch.empty(1, n, dtype=torch.bfloat16).pin_memory()
                xd = torch.empty(1, k, device=dev, dtype=torch.bfloat16)
                out["e2e_api"] = "RowShardedLinear.forward with explicit pinned-memory copies"

                def step_e2e():
                    for lin in layers:
                        xd.copy_(xh, non_blocking=True)
                        yh.copy_(lin(xd), non_blocking=True)
                    torch.cuda.synchronize()

            ms2 = time_gemv_set(step_e2e, copies, steps, warmup, dist)
            if world == 1:  # what arrived on the host is the device result, bit for bit
                xd = xh.to(dev)
                for (launch, yh_i), lin in zip(bound[:3], layers[:3]):
                    assert torch.equal(yh_i.to(dev), lin(xd)), "bind_host output differs from the device-resident forward"
            out["e2e_gbps"] = nbytes * copies * steps / (ms2 * 1e-3) / 1e9
            out["h2d"] = copies * k * 2
            out["d2h"] = copies * n * 2
        del layers
        torch.cuda.empty_cache()
        return out

    if args.profile_shape:  # ncu helper: just run one shape eagerly, no JSON contract
        r = bench_shape(args.profile_shape, args.profile_shape, args.steps, args.warmup, False)
        print(json.dumps(r))
        return
    def format_sweep():
        """BASELINE configs[3]: int4 / nf4 (any4, one global LUT) / any4 row-wise / mx4 x m in {1,4,8,16} at 4096^2,
        plus the A-layout int4 path (Int4Linear default), int8 and 16-bit weights at m = 1.  us per GEMM from a CUDA
        graph over `copies` distinct weight sets; informational (not part of the headline value)."""
        ops = torch.ops.tinygemm
        n = k = HEADLINE
        copies = max(3, int(2.6 * 126e6 / algorithmic_bytes(n, k)) + 1)
        gen = torch.Generator(device=dev).manual_seed(99)
        ws = [synth_layer(n, k, 500 + i, dev) for i in range(copies)]
        nf4 = torch.tensor([-1.0, -0.6962, -0.5251, -0.3949, -0.2844, -0.1848, -0.0911, 0.0, 0.0796, 0.1609, 0.2461,
                            0.3379, 0.4407, 0.5626, 0.723, 1.0], device=dev).bfloat16()
        sz32 = torch.stack([torch.rand(k // 32, n, generator=gen, device=dev) * 0.01 + 0.001,
                            torch.randn(k // 32, n, generator=gen, device=dev) * 0.01], 2).bfloat16().contiguous()
        exps = torch.randint(118, 130, (n, k // 32), generator=gen, device=dev, dtype=torch.int32).to(torch.uint8)

        def timed(fn):
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
            return round(e0.elapsed_time(e1) * 1e3 / (5 * copies), 2)

        out = {}
        for m in (1, 4, 8, 16):
            x = torch.randn(m, k, device=dev).bfloat16()
            out[f"m{m}"] = {
                "any4_rowwise_g128": timed(lambda: [ops.tinygemm_y_f16RM_x_f16RM_w_any4TC(x, w, G, sz, lut, True) for w, lut, sz in ws]),
                "nf4_global_g128": timed(lambda: [ops.tinygemm_y_f16RM_x_f16RM_w_any4TC(x, w, G, sz, nf4, True) for w, lut, sz in ws]),
                "int4_g128": timed(lambda: [ops.tinygemm_y_f16RM_x_f16RM_w_int4TC(x, w, G, sz, True) for w, lut, sz in ws]),
                "mx4_g32": timed(lambda: [ops.tinygemm_y_f16RM_x_f16RM_w_mx4TC(x, w, 32, exps, True) for w, lut, sz in ws]),
            }
        x = torch.randn(1, k, device=dev).bfloat16()
        wa = [w.view(n // 16, k // 64, 32, 4) for w, _, _ in ws]       # same bytes viewed as the A int4 layout (ik = 4)
        out["m1"]["int4_g128_A_layout(Int4Linear default)"] = timed(
            lambda: [ops.tinygemm_y_f16RM_x_f16RM_w_int4TC(w, x, G, sz, False) for w, (_, _, sz) in zip(wa, ws)])
        x16 = torch.randn(16, k, device=dev).bfloat16()  # (>= 3 rows: the weight is repacked into the B layout per call, then one pass)
        out["m16"]["int4_g128_A_layout(Int4Linear default)"] = timed(
            lambda: [ops.tinygemm_y_f16RM_x_f16RM_w_int4TC(w, x16, G, sz, False) for w, (_, _, sz) in zip(wa, ws)])
        # int8 and 16-bit weights (fragment-order kernel, untuned): 8 weight sets are enough to exceed L2 for 16-bit
        w8 = [torch.randint(-2**31, 2**31 - 1, (n // 8, k // 64, 32, 4), generator=gen, device=dev, dtype=torch.int64).to(torch.int32)
              for _ in range(16)]
        w16 = [torch.randn(n // 8, k // 32, 32, 8, generator=gen, device=dev).bfloat16() for _ in range(8)]
        sz0 = ws[0][2]
        copies_saved = copies
        copies = 16
        out["m1"]["int8_g128_B_layout"] = timed(lambda: [ops.tinygemm_y_f16RM_x_f16RM_w_int8TC(x, w, G, sz0, True) for w in w8])
        copies = 8
        out["m1"]["bf16_weights_B_layout"] = timed(lambda: [ops.tinygemm_y_f16RM_x_f16RM_w_f16TC(x, w, True) for w in w16])
        copies = copies_saved
        out["unit"] = "us per GEMM at n=k=4096, bf16, CUDA graph, weights rotated through > 2.5x L2"
        return out

    sampler = ClockSampler(local) if rank == 0 else None
    head = bench_shape(HEADLINE, HEADLINE, args.steps, args.warmup, True)
    extra = [bench_shape(s, s, max(3, args.steps // 2), args.warmup, False) for s in EXTRA_SHAPES]
    # the same headline measurement with the library's DEFAULT options (weights not declared static: the whole kernel,
    # not only its activation staging, is ordered behind the previous kernel of the stream)
    tgf.set_static_weights(False)
    head_default = bench_shape(HEADLINE, HEADLINE, max(3, args.steps // 2), args.warmup, False)
    tgf.set_static_weights(True)
    clocks = sampler.stop() if sampler else None
    sweep = format_sweep() if (rank == 0 and world == 1 and not args.no_sweep) else None

    # BASELINE.json's second metric (configs[2] / configs[4]): Llama-3-8B any4 g=128 single-token decode on this
    # many GPUs, one CUDA graph per token, row-sharded with the in-kernel exchange when N > 1 (bench_llama.py)
    llama = None
    if not args.no_llama:
        import bench_llama

        llama = {}
        for lm_head in ("bf16", "any4"):
            r = bench_llama.run_decode(dev, rank, world, dist, steps=args.llama_steps, warmup=3, exchange=args.exchange,
                                       lm_head_any4=lm_head == "any4", lib=lib)
            if rank == 0:
                llama["lm_head_" + lm_head] = {
                    "tok_s": round(r["value"], 1), "ms_per_token": round(r["ms_per_token"], 4),
                    "bytes_per_token_per_gpu": int(r["bytes_per_token_per_gpu"]),
                    "frac_of_hbm_bound": round(r["roofline"]["frac"], 4),
                    "library_launches_per_token": r["library_launches_per_token"]}
                llama.setdefault("workload", r["config"]["workload"])
                llama.setdefault("plumbing", r["config"]["plumbing"])
                llama.setdefault("parallelism", r["config"]["parallelism"])
        if rank == 0:
            llama["note"] = ("lm_head_bf16 = the reference's configuration (quantize.py:34-36 leaves lm_head unquantized, "
                             "cuBLAS GEMV); lm_head_any4 = lm_head through the same any4 kernel (SURVEY 8f-3)")

    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        cpu = cpu_baseline()

    if dist is not None:
        t = torch.tensor([1 if parity["ok"] else 0], device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MIN)
        parity["ok"] = bool(t.item())
        if not parity["ok"]:
            raise SystemExit("bench.py: the row-sharded output differs from the single-GPU output - nothing timed is valid")
    if rank == 0:
        nbytes = algorithmic_bytes(HEADLINE, HEADLINE)
        per_rank_bytes = nbytes / world  # each rank streams 1/world of the rows (x replicated, negligible)
        us = head["us_per_gemv"]
        achieved = per_rank_bytes / (us * 1e-6) / 1e9
        line = {
            "metric": METRIC, "value": head["gbps"], "unit": "GB/s", "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": head["ms"] / args.steps, "higher_is_better": True,
            "scaling": "strong", "vs_baseline": None, "dtype": "bf16", "data": "synthetic",
            "config": {
                "workload": WORKLOAD,
                "gemvs_per_step": head["copies"],
                "launch": "eager" if args.no_graph else "one CUDA graph per step (replayed K times)",
                "l2_policy": f"inputs larger than L2: {head['copies']} distinct weight sets = "
                             f"{head['copies'] * nbytes / 1e6:.0f} MB rotated every step",
                "parallelism": "1 GPU" if world == 1 else (
                    f"row-sharded x{world}, exchange inside the GEMV (tagged 8-byte words stored into every rank's symmetric buffer over NVLink, collected before the kernel exits; no barrier / collective launch)"
                    if args.exchange == "fused" else f"row-sharded x{world} + NCCL all-reduce on y"),
                "options": "TG_OPT_STATIC_WEIGHTS = 1 (weights / LUT / scales of a loaded model never change between launches)",
                "default_options": {"GBps": round(head_default["gbps"], 1), "us_per_gemv": round(head_default["us_per_gemv"], 3),
                                    "frac_of_peak": round(head_default["gbps"] / world / peak, 4)},
                "two_independent_chains(informational)": (
                    {"GBps": round(head["two_chains_gbps"], 1), "us_per_gemv": round(head["two_chains_us_per_gemv"], 3),
                     "frac_of_peak": round(head["two_chains_gbps"] / peak, 4),
                     "what": "the same GEMVs as two stream-ordered chains on two streams inside one CUDA graph"}
                    if "two_chains_gbps" in head else None),
                "other_shapes": {f"{e['n']}x{e['k']}": {"GBps": round(e["gbps"], 1), "us_per_gemv": round(e["us_per_gemv"], 3),
                                                         "frac_of_peak": round(e["gbps"] / world / peak, 4)} for e in extra},
            },
            "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                         "traffic": ncu_traffic() if world == 1 else None, "peak_source": peak_src + " (of measured)",
                         "kernel": "gemv_w4_b_kernel<bf16, ik=4, m=1>", "us_per_launch": us,
                         "algorithmic_bytes_per_launch": per_rank_bytes},
            "e2e": {"value": head["e2e_gbps"], "unit": "GB/s", "h2d_bytes_per_step": head["h2d"],
                    "d2h_bytes_per_step": head["d2h"], "api": head.get("e2e_api"),
                    "eager_launches_GBps": (round(head["e2e_eager_gbps"], 1) if "e2e_eager_gbps" in head else None)},
            "gpu_launches": head["launches"],
            "parity_checked": (parity["ok"] if parity["checked"] else None),
            "parity_frac_bit_equal": parity.get("frac_bit_equal"),
            "clocks": clocks,
        }
        if cpu is not None:
            line["cpu_baseline"] = cpu
        if sweep is not None:
            line["config"]["format_sweep_us"] = sweep
        if llama is not None:
            line["config"]["llama_decode"] = llama
        print(json.dumps(line))
    if dist is not None:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-graph", action="store_true", help="time eager launches instead of CUDA-graph replays")
    ap.add_argument("--exchange", default="fused", choices=["fused", "nccl"],
                    help="N > 1: 'fused' = GEMV epilogue stores into all ranks' symmetric buffers + one barrier; "
                         "'nccl' = zero-padded all-reduce")
    ap.add_argument("--no-sweep", action="store_true", help="skip the informational format / m sweep")
    ap.add_argument("--no-llama", action="store_true", help="skip the Llama-3-8B decode measurement")
    ap.add_argument("--llama-steps", type=int, default=30, help="timed decode steps (CUDA-graph replays) per variant")
    ap.add_argument("--profile-shape", type=int, default=0, help="(for ncu) run only the n=k=N GEMV set")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "ours" else args.warmup
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
