"""Small driver for compute-sanitizer runs: a few instances through every kernel path."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from cave_b200 import cave_forward_backward, synth

dev = torch.device("cuda:0")
for kind, B in (("tsp20", 6), ("sp5", 4), ("vrp20", 4)):
    insts = synth.make_batch(kind, B, seed=1)
    A = synth.densify(insts, device=dev)
    pred = torch.tensor(synth.predictions(insts, 1, "near"), device=dev)
    for mode in (0, 1, 2):
        for prec in ("fp64", "fp32"):
            out = cave_forward_backward(pred, A, -1.0, mode, 0.2, "mean", precision=prec, want_proj=mode != 2, want_status=True)
torch.manual_seed(0)
A = torch.randn(4, 15, 10, device=dev); c = torch.randn(4, 10, device=dev)
out = cave_forward_backward(c, A, -1.0, 0, want_proj=True, want_status=True)          # Lawson-Hanson path
A = torch.randn(2, 40, 3000, device=dev); c = torch.randn(2, 3000, device=dev)
out = cave_forward_backward(c, A, -1.0, 0, want_proj=True, want_status=True)          # tile scan kernel, CSR fallback
os.environ["CAVE_SCAN_KERNEL"] = "tile"
insts = synth.make_batch("tsp20", 3, seed=2)
out = cave_forward_backward(torch.tensor(synth.predictions(insts, 2, "near"), device=dev), synth.densify(insts, device=dev), -1.0, 1)
torch.cuda.synchronize()
print("sanitize case done", float(out["loss"]))
