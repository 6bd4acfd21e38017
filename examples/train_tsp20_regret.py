#!/usr/bin/env python
"""
BASELINE.json configs[1] end to end: TSP-20 DFJ (190 edges), linear predictor, batch 32, CaVE Exact / CaVE+ / CaVE Hybrid
with solver='cuda' on a device-resident packed dataset, normalised decision regret on a held-out test set with EXACT
tours (Held-Karp on the GPU, cave_b200/tsp_exact.py) — next to the 2-stage MSE baseline and to an emulation of the
reference's default backend (Clarabel truncated at max_iter = 3: examples/clarabel_emulation.py, parity unpinned).
Template: /root/reference/code_sample.py:14-61.  No Gurobi / PyEPO / cvxpy.  The slides report for TSP-20,
degree 4: CaVE-E 7.35 +- 0.40 %, CaVE+ 6.20 +- 0.24 %, CaVE-H 7.69 +- 0.33 % (dataset size and epochs not stated).

    python examples/train_tsp20_regret.py [--seeds 3] [--epochs 10] [--train 1000] [--test 1000]
"""
import argparse
import json
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))

import numpy as np  # noqa: E402
import torch  # noqa: E402
from torch import nn  # noqa: E402

from cave_b200 import EPO, exactConeAlignedCosine, innerConeAlignedCosine, pack_constraints, synth, tsp_exact  # noqa: E402
from clarabel_emulation import ClarabelTruncatedCosine  # noqa: E402

N_NODES = 20


class Model:
    modelSense = EPO.MINIMIZE


def run(method, data, epochs=10, batch=32, seed=0, lr=1e-2):
    torch.manual_seed(seed)
    dev = torch.device("cuda", torch.cuda.current_device())
    X, C, A, pack, xte, cte, obj_te = data
    reg = nn.Linear(X.shape[1], C.shape[1]).to(dev)
    opt = torch.optim.Adam(reg.parameters(), lr=lr)
    if method == "cave-e":
        loss_fn = exactConeAlignedCosine(Model(), solver="cuda")
    elif method == "cave+":
        loss_fn = innerConeAlignedCosine(Model(), solver="cuda", inner_ratio=0.2, seed=seed)
    elif method == "cave-h":
        loss_fn = innerConeAlignedCosine(Model(), solver="cuda", inner_ratio=0.2, solve_ratio=0.3, seed=seed)
    elif method == "clarabel3":
        loss_fn = ClarabelTruncatedCosine(minimize=True, max_iter=3)
    else:
        loss_fn = None      # 2-stage MSE
    n_train = X.shape[0]
    t0 = time.perf_counter()
    last = []
    for epoch in range(epochs):
        perm = torch.randperm(n_train, device=dev)
        for s in range(0, n_train, batch):
            idx = perm[s:s + batch]
            cp = reg(X[idx])
            if loss_fn is None:
                loss = ((cp - C[idx]) ** 2).mean()
            elif method == "clarabel3":
                loss = loss_fn(cp, A[idx])
            else:
                loss = loss_fn(cp, pack, index=idx.to(torch.int32))
            opt.zero_grad()
            loss.backward()
            opt.step()
            if epoch == epochs - 1:
                last.append(float(loss.detach()))
    torch.cuda.synchronize()
    train_s = time.perf_counter() - t0
    pred = reg(torch.tensor(xte, device=dev)).detach().cpu().numpy()
    return tsp_exact.normalised_regret(pred, cte, N_NODES, true_obj=obj_te), float(np.mean(last)), train_s


def side_by_side_loss(data, n=256):
    """Loss of the same predictions under the exact push-inside target (solver='cuda', the reference's nnls branch) and under
    the emulated truncated-Clarabel target: north_star asks for the two side by side."""
    dev = torch.device("cuda", torch.cuda.current_device())
    X, C, A, pack, *_ = data
    torch.manual_seed(0)
    pred = C[:n] * (1.0 + 0.5 * torch.randn_like(C[:n]))
    idx = torch.arange(n, device=dev, dtype=torch.int32)
    l_cuda = innerConeAlignedCosine(Model(), solver="cuda", inner_ratio=0.2, seed=0, reduction="none")(pred, pack, index=idx)
    l_exact = exactConeAlignedCosine(Model(), solver="cuda", reduction="none")(pred, pack, index=idx)
    l_emu = ClarabelTruncatedCosine(max_iter=3, reduction="none")(pred, A[:n])
    l_emu50 = ClarabelTruncatedCosine(max_iter=25, reduction="none")(pred, A[:n])
    return {"n": n, "loss_cave_plus_cuda_mean": float(l_cuda.mean()), "loss_exact_cuda_mean": float(l_exact.mean()),
            "loss_clarabel3_emulation_mean": float(l_emu.mean()), "loss_ipm25_emulation_mean": float(l_emu50.mean()),
            "max_abs_diff_ipm25_vs_exact": float((l_emu50 - l_exact).abs().max())}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--seeds", type=int, default=3)
    ap.add_argument("--epochs", type=int, default=10)
    ap.add_argument("--train", type=int, default=1000)
    ap.add_argument("--test", type=int, default=1000)
    ap.add_argument("--feat", type=int, default=10)
    ap.add_argument("--deg", type=int, default=4)
    ap.add_argument("--methods", default="2stage,cave-e,cave+,cave-h,clarabel3")
    ap.add_argument("--json", default="")
    a = ap.parse_args()
    dev = torch.device("cuda", torch.cuda.current_device())
    results = {}
    sbs = None
    for seed in range(a.seeds):
        x, c = tsp_exact.gen_data(a.train + a.test, a.feat, N_NODES, a.deg, 0.5, seed=135 + seed)
        xtr, ctr, xte, cte = x[:a.train], c[:a.train], x[a.train:], c[a.train:]
        t0 = time.perf_counter()
        _, _, tours = tsp_exact.solve(ctr, N_NODES)
        _, obj_te, _ = tsp_exact.solve(cte, N_NODES)
        t_solve = time.perf_counter() - t0
        rng = np.random.default_rng(seed)
        insts = [tsp_exact.binding_constraints(t, rng, 8) for t in tours]
        A = synth.densify(insts, device=dev)
        pack = pack_constraints(A)
        data = (torch.tensor(xtr, device=dev), torch.tensor(ctr, device=dev), A, pack, xte, cte, obj_te)
        if seed == 0:
            sbs = side_by_side_loss(data)
            print("side by side (same predictions):", json.dumps(sbs), flush=True)
        for m in a.methods.split(","):
            r, l, ts = run(m, data, a.epochs, 32, seed)
            results.setdefault(m, []).append(r)
            print(f"seed {seed} {m:10s}: normalised test regret {100 * r:6.2f} %   last-epoch loss {l:.5f}   train {ts:6.1f} s   "
                  f"(Held-Karp for {a.train + a.test} instances: {t_solve:.2f} s)", flush=True)
    summary = {m: {"mean_pct": 100 * float(np.mean(v)), "std_pct": 100 * float(np.std(v)), "runs_pct": [100 * float(t) for t in v]}
               for m, v in results.items()}
    print(json.dumps({"config": vars(a), "regret": summary, "side_by_side_loss": sbs}))
    if a.json:
        json.dump({"config": vars(a), "regret": summary, "side_by_side_loss": sbs}, open(a.json, "w"), indent=1)


if __name__ == "__main__":
    main()
