"""2-GPU debug of the in-kernel exchange: where do sharded and single-GPU outputs differ?  (torchrun, 2 ranks)"""
import os, sys
import torch
import torch.distributed as dist
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
from any4_b200.modules import Any4Linear, RowShardedLinear
from bench import synth_layer, G

rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
torch.cuda.set_device(rank)
dev = torch.device("cuda", rank)
dist.init_process_group("nccl", device_id=dev)
for n, k in [(4096, 4096), (1024, 2048)]:
    lin = Any4Linear(k, n, bias=False, device=dev, dtype=torch.bfloat16, group_size=G)
    w, lut, sz = synth_layer(n, k, 1234, dev)
    lin.weight.data, lin.lut.data, lin.scales_and_zeros.data = w, lut, sz
    lin.weight_reshaped = True
    sh = RowShardedLinear(lin, rank, world, fused=True, max_features=11008)
    for m in (1, 2, 5):
        x = torch.randn(m, k, device=dev, generator=torch.Generator(device=dev).manual_seed(5 + m)).bfloat16()
        for rep in range(2):
            got = sh(x).float().clone()
            want = lin(x).float()
            torch.cuda.synchronize()
            bad = (got != want)
            rel = ((got - want).abs() / (want.abs() + 1e-6))
            print(f"[rank {rank}] n={n} k={k} m={m} rep={rep}: mismatches {int(bad.sum())}/{bad.numel()}, max rel {float(rel.max()):.3g}, "
                  f"first bad cols {bad.nonzero()[:6].tolist()}, zeros in got {int((got == 0).sum())}", flush=True)
dist.barrier()
dist.destroy_process_group()
