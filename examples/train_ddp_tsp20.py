#!/usr/bin/env python
"""
Multi-GPU training with the CaVE+ loss: one process per GPU, the dataset sharded by instance, the
projection/loss computed locally with NO collective; only the predictor's gradients are all-reduced (DDP over
NCCL) and one 2-element all-reduce logs the global loss.

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 examples/train_ddp_tsp20.py
"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402
from torch import nn  # noqa: E402
from torch.nn.parallel import DistributedDataParallel as DDP  # noqa: E402

from cave_b200 import EPO, innerConeAlignedCosine, pack_constraints  # noqa: E402
from cave_b200.parallel import global_mean_loss, instance_shard  # noqa: E402
from examples.train_tsp20_cave_plus import make_dataset  # noqa: E402


class Model:
    modelSense = EPO.MINIMIZE


def main(num_data=512, epochs=5, batch_per_rank=32):
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist.init_process_group("nccl", device_id=dev)
    torch.manual_seed(0)                                           # identical initial predictor on every rank
    feats, costs, ctrs = make_dataset(num_data, device=dev)        # same synthetic dataset everywhere ...
    lo, hi = instance_shard(num_data, rank, world)                 # ... each rank keeps only its shard resident
    feats, ctrs = feats[lo:hi], ctrs[lo:hi].contiguous()
    pack = pack_constraints(ctrs, keep_dense=False)
    del ctrs
    reg = DDP(nn.Linear(feats.shape[1], costs.shape[1]).to(dev), device_ids=[local])
    cave = innerConeAlignedCosine(Model(), solver="cuda", inner_ratio=0.2, seed=0)      # reduction='mean' per shard
    opt = torch.optim.Adam(reg.parameters(), lr=1e-2)
    n_local = hi - lo
    for epoch in range(epochs):
        perm = torch.randperm(n_local, device=dev)
        tot = torch.zeros((), device=dev, dtype=torch.float64)
        for s in range(0, n_local - batch_per_rank + 1, batch_per_rank):     # equal shards -> DDP mean == global mean
            idx = perm[s:s + batch_per_rank]
            loss = cave(reg(feats[idx]), pack, index=idx.to(torch.int32))
            opt.zero_grad()
            loss.backward()                                         # DDP all-reduces the 10x190 predictor here
            opt.step()
            tot += loss.detach().double() * len(idx)
        g = global_mean_loss(tot, (n_local // batch_per_rank) * batch_per_rank)
        w = torch.cat([p.detach().flatten() for p in reg.parameters()])
        wmax, wmin = w.clone(), w.clone()
        dist.all_reduce(wmax, op=dist.ReduceOp.MAX)
        dist.all_reduce(wmin, op=dist.ReduceOp.MIN)
        if rank == 0:
            print(f"epoch {epoch}: global mean loss {g.item():.5f}   max parameter drift across ranks "
                  f"{float((wmax - wmin).abs().max()):.1e}")
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
