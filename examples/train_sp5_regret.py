#!/usr/bin/env python
"""
BASELINE.json configs[0] end to end: shortest path on the 5x5 grid (40 arcs), linear predictor, batch 32,
CaVE Exact / CaVE+ / CaVE Hybrid with solver='cuda', normalised regret on a held-out test set — next to the
2-stage MSE baseline.  No Gurobi / PyEPO: exact paths by DP, binding constraints from the vertex
(cave_b200/sp_grid.py).  The slides report 8-11 % regret for this problem at degree 4.
"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

import numpy as np  # noqa: E402
import torch  # noqa: E402
from torch import nn  # noqa: E402

from cave_b200 import EPO, exactConeAlignedCosine, innerConeAlignedCosine, pack_constraints, sp_grid, synth  # noqa: E402


class Model:
    modelSense = EPO.MINIMIZE


def run(method="cave+", n_train=1000, n_test=1000, num_feat=5, deg=4, epochs=10, batch=32, seed=0, verbose=True):
    torch.manual_seed(seed)
    dev = torch.device("cuda", torch.cuda.current_device())
    x, c = sp_grid.gen_data(n_train + n_test, num_feat, 5, deg, 0.5, seed=135 + seed)
    xtr, ctr, xte, cte = x[:n_train], c[:n_train], x[n_train:], c[n_train:]
    sols, _ = sp_grid.solve(ctr)
    insts = [sp_grid.binding_constraints(s) for s in sols]
    pack = pack_constraints(synth.densify(insts, device=dev), keep_dense=False)
    X, C = torch.tensor(xtr, device=dev), torch.tensor(ctr, device=dev)
    reg = nn.Linear(num_feat, c.shape[1]).to(dev)
    opt = torch.optim.Adam(reg.parameters(), lr=1e-2)
    if method == "cave-e":
        loss_fn = exactConeAlignedCosine(Model(), solver="cuda")
    elif method == "cave+":
        loss_fn = innerConeAlignedCosine(Model(), solver="cuda", inner_ratio=0.2, seed=seed)
    elif method == "cave-h":
        loss_fn = innerConeAlignedCosine(Model(), solver="cuda", inner_ratio=0.2, solve_ratio=0.3, seed=seed)
    else:
        loss_fn = None      # 2-stage MSE
    regret0 = sp_grid.normalised_regret(reg(torch.tensor(xte, device=dev)).detach().cpu().numpy(), cte)
    for epoch in range(epochs):
        perm = torch.randperm(n_train, device=dev)
        for s in range(0, n_train, batch):
            idx = perm[s:s + batch]
            cp = reg(X[idx])
            loss = loss_fn(cp, pack, index=idx.to(torch.int32)) if loss_fn is not None else ((cp - C[idx]) ** 2).mean()
            opt.zero_grad()
            loss.backward()
            opt.step()
        if verbose:
            r = sp_grid.normalised_regret(reg(torch.tensor(xte, device=dev)).detach().cpu().numpy(), cte)
            print(f"{method:7s} epoch {epoch:3d} loss {loss.item():9.5f} test regret {100 * r:6.2f} %")
    regret = sp_grid.normalised_regret(reg(torch.tensor(xte, device=dev)).detach().cpu().numpy(), cte)
    return regret0, regret


if __name__ == "__main__":
    for m in ("2stage", "cave-e", "cave+", "cave-h"):
        r0, r1 = run(m, verbose=False)
        print(f"{m:7s}: normalised test regret {100 * r0:6.2f} % (untrained) -> {100 * r1:6.2f} % after 10 epochs")
