"""The SP 5x5 instance (seed 1001, uniform regime, #3905) on which the structured Newton solver hits its iteration cap."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from cave_b200 import cave_forward_backward, synth
from oracle import cave_oracle as O
dev = torch.device("cuda:0")
insts = synth.make_batch("sp5", 4096, seed=1001)
A = synth.densify(insts, device=dev)
pred = torch.tensor(synth.predictions(insts, 1001, "uniform"), device=dev)
i = 3905
for prec in ("fp64", "fp32"):
    for kw in ({}, {"max_iter": 2000}):
        out = cave_forward_backward(pred[i:i+1], A[i:i+1], -1.0, 0, 0.2, "none", precision=prec, want_proj=True, want_status=True, **kw)
        p64, r64 = O.batch_project((-pred[i:i+1]).double().cpu().numpy(), A[i:i+1].cpu().numpy(), fp64=True)
        err = np.abs(out["proj"].double().cpu().numpy() - p64).max() / max(np.abs(p64).max(), 1e-30)
        print(prec, kw, "status", hex(int(out["status"][0])), "iters", int(out["iters"][0]), "rnorm", float(out["rnorm"][0]), "oracle rnorm", float(r64[0]),
              "proj rel err", err, flush=True)
ins = insts[i]
print("m", ins.m, "rows nnz histogram", np.bincount(np.bincount(ins.rows, minlength=ins.m)).tolist())
np.savez_compressed("gpurun_out/sp5_1001_3905.npz", A=A[i].cpu().numpy(), pred=pred[i].cpu().numpy())
