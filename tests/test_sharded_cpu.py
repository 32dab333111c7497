"""CPU, world_size = 2 over gloo: the host logic of the row-sharded multi-GPU path (shard slicing,
zero-padded buffers, ONE all-reduce, bias after the reduction).  The local GEMV is replaced by the
oracle's dense product here - there is no GPU in this test - so this checks the N > 1 plumbing, not
the kernel."""
import os
import socket

import pytest
import torch
import torch.multiprocessing as mp


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, out):
    import torch.distributed as dist

    from any4_b200.modules import Any4Linear, RowShardedLinear
    from oracle import dequant, layouts

    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    gen = torch.Generator().manual_seed(0)
    n, k, g = 64, 256, 64
    codes = torch.randint(0, 16, (n, k), generator=gen, dtype=torch.int32)
    lut = ((torch.rand(n, 16, generator=gen) * 15).sort(1).values.bfloat16() - 8)
    sz = torch.stack([torch.rand(k // g, n, generator=gen) * 0.01 + 0.001, torch.randn(k // g, n, generator=gen) * 0.01], 2).bfloat16()
    bias = torch.randn(n, generator=gen).bfloat16()
    x = torch.randn(3, k, generator=gen).bfloat16()
    full = Any4Linear(k, n, bias=True, dtype=torch.bfloat16, group_size=g)
    full.weight.data = torch.from_numpy(layouts.to_Bint4(codes.numpy(), 4))
    full.lut.data, full.scales_and_zeros.data, full.bias.data = lut, sz, bias
    full.weight_reshaped = True
    sh = RowShardedLinear(full, rank, world)
    assert sh.local.weight.shape[0] == n // 8 // world

    def local_gemm(x2d):  # oracle stand-in for the CUDA kernel on this GPU-less box
        c = torch.from_numpy(layouts.from_Bint4(sh.local.weight.data.numpy()))
        w = dequant.dequant_lut(c, sh.local.lut.data, sh.local.scales_and_zeros.data, g, torch.bfloat16)
        return dequant.gemm(x2d, w)

    sh.local._gemm = local_gemm
    y = sh(x)
    w = dequant.dequant_lut(codes, lut, sz, g, torch.bfloat16)
    want = dequant.gemm(x, w) + bias
    ok = torch.equal(y, want)
    if rank == 0:
        out.put(ok)
    dist.barrier()
    dist.destroy_process_group()


def _worker_fused(rank, world, port, out):
    """q|k|v-style row fusion (modules.fuse_rows) of two layers, then row-sharded over the ranks: every rank owns a
    contiguous slice of the CONCATENATED rows, the all-reduce reassembles [layer0 | layer1]."""
    import torch.distributed as dist

    from any4_b200.modules import Any4Linear, RowShardedLinear, fuse_rows
    from oracle import dequant, layouts

    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    gen = torch.Generator().manual_seed(1)
    k, g = 256, 64
    x = torch.randn(2, k, generator=gen).bfloat16()
    lins, wants = [], []
    for n in (48, 16):
        codes = torch.randint(0, 16, (n, k), generator=gen, dtype=torch.int32)
        lut = ((torch.rand(n, 16, generator=gen) * 15).sort(1).values.bfloat16() - 8)
        sz = torch.stack([torch.rand(k // g, n, generator=gen) * 0.01 + 0.001, torch.randn(k // g, n, generator=gen) * 0.01], 2).bfloat16()
        lin = Any4Linear(k, n, bias=False, dtype=torch.bfloat16, group_size=g)
        lin.weight.data = torch.from_numpy(layouts.to_Bint4(codes.numpy(), 4))
        lin.lut.data, lin.scales_and_zeros.data = lut, sz
        lin.weight_reshaped = True
        lins.append(lin)
        wants.append(dequant.gemm(x, dequant.dequant_lut(codes, lut, sz, g, torch.bfloat16)))
    sh = RowShardedLinear(fuse_rows(lins), rank, world)
    assert sh.local.weight.shape[0] == 64 // 8 // world

    def local_gemm(x2d):
        c = torch.from_numpy(layouts.from_Bint4(sh.local.weight.data.numpy()))
        return dequant.gemm(x2d, dequant.dequant_lut(c, sh.local.lut.data, sh.local.scales_and_zeros.data, g, torch.bfloat16))

    sh.local._gemm = local_gemm
    ok = torch.equal(sh(x), torch.cat(wants, -1))
    if rank == 0:
        out.put(ok)
    dist.barrier()
    dist.destroy_process_group()


def test_fused_rows_sharded_gloo_world2():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker_fused, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    ok = q.get(timeout=120)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert ok


def test_row_sharded_allreduce_gloo_world2():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    ok = q.get(timeout=120)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert ok


def test_shard_rejects_partial_tiles():
    from any4_b200.modules import Any4Linear, RowShardedLinear

    m = Any4Linear(256, 24, bias=False, dtype=torch.bfloat16)
    m.weight.data = torch.zeros(3, 4, 32, 2, dtype=torch.int32)
    m.weight_reshaped = True
    with pytest.raises(ValueError):
        RowShardedLinear(m, 0, 2)
