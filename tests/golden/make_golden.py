"""
Generates tests/golden/*.npz by running the UNMODIFIED reference
(/root/reference/src/cave.py, solver='nnls') in the build container.

    python tests/golden/make_golden.py

/root/reference does not exist on the GPU box, so the vectors are committed.  PyEPO is not
installed in this image; tests/golden/_pyepo_stub provides the three symbols the reference
imports (see its docstring).  Inputs come from fixed seeds (torch.manual_seed as in
test/test_func.py:34-43 and 280-283) and from cave_b200.synth.
"""
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, os.path.join(HERE, "_pyepo_stub"))
sys.path.insert(0, "/root/reference")
sys.path.insert(0, ROOT)

from pyepo import EPO  # noqa: E402
from src.cave import _batch_project, exactConeAlignedCosine, innerConeAlignedCosine  # noqa: E402

from cave_b200 import synth  # noqa: E402


class _Model:
    def __init__(self, sense):
        self.modelSense = sense


def run(name, pred, ctrs, cls="exact", sense="min", reduction="mean", **kw):
    model = _Model(EPO.MINIMIZE if sense == "min" else EPO.MAXIMIZE)
    pred = pred.clone().requires_grad_(True)
    if cls == "exact":
        mod = exactConeAlignedCosine(model, solver="nnls", reduction=reduction)
        mode = 0
    else:
        mod = innerConeAlignedCosine(model, solver="nnls", reduction=reduction, **kw)
        mode = 2 if kw.get("solve_ratio", 1.0) == 0 else 1
    loss = mod(pred, ctrs)
    loss.sum().backward()
    sign = -1.0 if sense == "min" else 1.0
    with torch.no_grad():
        proj, rnorm = _batch_project(sign * pred.detach(), ctrs, "nnls", None, 1, None)
    out = dict(pred=pred.detach().numpy(), ctrs=ctrs.numpy(), loss=loss.detach().numpy(),
               grad=pred.grad.numpy(), proj=proj.numpy(), rnorm=rnorm.numpy(),
               mode=np.int32(mode), minimize=np.bool_(sense == "min"),
               inner_ratio=np.float64(kw.get("inner_ratio", 0.2)), reduction=np.str_(reduction))
    np.savez_compressed(os.path.join(HERE, name + ".npz"), **out)
    print(f"{name}: loss={np.asarray(out['loss']).ravel()[:3]} |grad|={np.abs(out['grad']).max():.3e}")


def structured(kind, batch, seed, regime):
    insts = synth.make_batch(kind, batch, seed)
    ctrs = synth.densify(insts)
    pred = torch.from_numpy(synth.predictions(insts, seed, regime))
    return pred, ctrs


if __name__ == "__main__":
    # (1) test/test_func.py:34-43 data: rand >= 0, MINIMIZE -> lambda = 0, loss == 1, grad == 0
    torch.manual_seed(0)
    pred, ctrs = torch.rand(32, 10), torch.rand(32, 15, 10)
    run("tf_rand_exact", pred, ctrs)
    run("tf_rand_inner", pred, ctrs, cls="inner", seed=42)
    run("tf_rand_heur", pred, ctrs, cls="inner", solve_ratio=0, seed=42)
    # (12) test/test_func.py:280-283 data: randn, seed 1, B=8, m=15, d=10
    torch.manual_seed(1)
    pred, ctrs = torch.randn(8, 10), torch.randn(8, 15, 10)
    run("tf_randn_exact", pred, ctrs)
    run("tf_randn_exact_max_none", pred, ctrs, sense="max", reduction="none")
    run("tf_randn_inner_sum", pred, ctrs, cls="inner", reduction="sum", seed=3, inner_ratio=0.35)
    run("tf_randn_heur", pred, ctrs, cls="inner", solve_ratio=0, seed=3)
    # (2) zero prediction (test/test_func.py:182-193)
    torch.manual_seed(0)
    run("zero_pred_exact", torch.zeros(2, 6), torch.rand(2, 3, 6))
    # (3)(4) padded zero rows and an empty instance
    torch.manual_seed(2)
    pred, ctrs = torch.randn(4, 6), torch.randn(4, 5, 6)
    ctrs = torch.cat([ctrs, torch.zeros(4, 7, 6)], dim=1)
    ctrs[3] = 0.0
    run("padded_empty_inner", pred, ctrs, cls="inner", seed=0, reduction="none")
    # (5) prediction inside the cone -> rnorm ~ 0, un-pushed target
    torch.manual_seed(3)
    ctrs = torch.randn(4, 5, 8)
    pred = -(torch.rand(4, 5).unsqueeze(2) * ctrs).sum(dim=1)
    run("inside_inner", pred, ctrs, cls="inner", seed=0, reduction="none")
    # config 1: shortest path 5x5, exact, batch 32
    for regime in ("uniform", "near"):
        pred, ctrs = structured("sp5", 32, 0, regime)
        run(f"sp5_exact_{regime}", pred, ctrs)
    # config 2: TSP-20, CaVE+ inner_ratio 0.2
    for regime in ("uniform", "near"):
        pred, ctrs = structured("tsp20", 8, 0, regime)
        run(f"tsp20_inner_{regime}", pred, ctrs, cls="inner", seed=0, reduction="none")
    # config 4: VRP-20 ragged
    pred, ctrs = structured("vrp20", 8, 0, "near")
    run("vrp20_inner_near", pred, ctrs, cls="inner", seed=0, reduction="none")
    run("vrp20_heur_near", pred, ctrs, cls="inner", solve_ratio=0, seed=0, reduction="none")
    # config 3 shape, two instances (projection only is expensive on the CPU: ~1.3 s each)
    pred, ctrs = structured("tsp50", 2, 0, "near")
    run("tsp50_inner_near", pred, ctrs, cls="inner", seed=0, reduction="none")
