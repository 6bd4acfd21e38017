"""N > 1 path on CPU: world_size-2 gloo processes shard a batch by instance, compute their shard's loss
and gradient (the per-shard compute is the oracle here: no GPU in this container), and the sharded
results reassemble to the single-process answer with no data-path collective."""
import os
import socket
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, out_dir):
    sys.path.insert(0, ROOT)
    os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from cave_b200 import synth
    from cave_b200.parallel import global_mean_loss, instance_shard, sharded_grad_scale
    from oracle import cave_oracle as O
    B = 11                                              # ragged split: 6 + 5
    insts = synth.make_batch("sp5", B, seed=3)
    ctrs, pred = synth.densify(insts).numpy(), synth.predictions(insts, 3, "near")
    lo, hi = instance_shard(B, rank, world)
    out = O.forward_backward(pred[lo:hi], ctrs[lo:hi], mode=1, reduction="mean", fp64=True)
    # what DDP would do with a linear predictor's gradient: average over ranks after ragged rescaling
    g = torch.zeros(B, pred.shape[1], dtype=torch.float64)
    g[lo:hi] = torch.from_numpy(out["grad"]) * sharded_grad_scale(hi - lo, B, world)
    dist.all_reduce(g)
    g /= world
    loss = global_mean_loss(torch.tensor(out["loss_i"].sum()), hi - lo)
    if rank == 0:
        np.savez(os.path.join(out_dir, "out.npz"), grad=g.numpy(), loss=float(loss))
    dist.destroy_process_group()


def test_two_rank_instance_sharding_matches_single_process(tmp_path):
    from cave_b200 import synth
    from oracle import cave_oracle as O
    port = _free_port()
    mp.spawn(_worker, args=(2, port, str(tmp_path)), nprocs=2, join=True)
    z = np.load(tmp_path / "out.npz")
    insts = synth.make_batch("sp5", 11, seed=3)
    ref = O.forward_backward(synth.predictions(insts, 3, "near"), synth.densify(insts).numpy(), mode=1,
                             reduction="mean", fp64=True)
    np.testing.assert_allclose(z["loss"], ref["loss"], rtol=1e-12)
    np.testing.assert_allclose(z["grad"], ref["grad"], rtol=1e-12, atol=1e-15)


def test_instance_shard_covers_batch_exactly():
    from cave_b200.parallel import instance_shard
    for B in (0, 1, 7, 8, 4096, 4099):
        for world in (1, 2, 3, 8):
            cuts = [instance_shard(B, r, world) for r in range(world)]
            assert cuts[0][0] == 0 and cuts[-1][1] == B
            assert all(cuts[i][1] == cuts[i + 1][0] for i in range(world - 1))
            sizes = [hi - lo for lo, hi in cuts]
            assert max(sizes) - min(sizes) <= 1
    with pytest.raises(ValueError):
        instance_shard(8, 2, 2)
