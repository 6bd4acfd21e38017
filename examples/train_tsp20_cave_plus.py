#!/usr/bin/env python
"""
Counterpart of the reference's code_sample.py (code_sample.py:14-61) with solver='cuda': a linear predictor
10 -> 190 trained with CaVE+ (inner_ratio 0.2) on TSP-20 binding constraints, batch 32, Adam 1e-2.

No Gurobi/PyEPO here, so the dataset is synthetic (cave_b200.synth): random tours and
their binding constraints; costs are generated so that the stored tour is (near) optimal, and features are
a noisy linear image of the costs.  The constraints are packed ONCE on the device (CavePack) and batches are
addressed by instance index — the device-resident replacement of DataLoader + collate_fn.
"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

import numpy as np  # noqa: E402
import torch  # noqa: E402
from torch import nn  # noqa: E402

from cave_b200 import EPO, innerConeAlignedCosine, pack_constraints, synth  # noqa: E402


class Model:
    modelSense = EPO.MINIMIZE


def make_dataset(n, num_feat=10, seed=42, device="cuda"):
    insts = synth.make_batch("tsp20", n, seed=seed)
    costs = synth.predictions(insts, seed, "near")                       # [n, 190]
    rng = np.random.default_rng(seed)
    proj = rng.standard_normal((costs.shape[1], num_feat)).astype(np.float32) / np.sqrt(costs.shape[1])
    feats = (costs @ proj + 0.05 * rng.standard_normal((n, num_feat))).astype(np.float32)
    ctrs = synth.densify(insts, device=device)
    return torch.tensor(feats, device=device), torch.tensor(costs, device=device), ctrs


def train(num_data=256, epochs=10, batch=32, seed=0, verbose=True):
    torch.manual_seed(seed)
    dev = torch.device("cuda", torch.cuda.current_device())
    feats, costs, ctrs = make_dataset(num_data, device=dev)
    pack = pack_constraints(ctrs, keep_dense=False)                      # constraints live on the device from here on
    del ctrs
    reg = nn.Linear(feats.shape[1], costs.shape[1]).to(dev)
    cave = innerConeAlignedCosine(Model(), solver="cuda", inner_ratio=0.2, seed=seed)
    opt = torch.optim.Adam(reg.parameters(), lr=1e-2)
    history = []
    for epoch in range(epochs):
        perm = torch.randperm(num_data, device=dev)
        tot = 0.0
        for s in range(0, num_data, batch):
            idx = perm[s:s + batch].to(torch.int32)
            loss = cave(reg(feats[idx.long()]), pack, index=idx)
            opt.zero_grad()
            loss.backward()
            opt.step()
            tot += loss.item() * len(idx)
        history.append(tot / num_data)
        if verbose:
            print(f"Epoch {epoch:4d}, Loss: {history[-1]:8.4f}")
    return history


if __name__ == "__main__":
    train()
