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
# round 2 paths: warm pack with the cached solver setup, sparse ingestion, the dense (tensor-core Gram) path, Held-Karp
from cave_b200 import SparseConstraints, pack_constraints, pack_constraints_sparse, tsp_exact
os.environ.pop("CAVE_SCAN_KERNEL", None)
insts = synth.make_batch("tsp20", 4, seed=3)
A = synth.densify(insts, device=dev)
pred = torch.tensor(synth.predictions(insts, 3, "near"), device=dev)
pk = pack_constraints(A)
out = cave_forward_backward(pred, A, -1.0, 1, pack=pk, want_status=True)
pks = pack_constraints_sparse(SparseConstraints.from_instances(insts))
out = cave_forward_backward(pred, None, -1.0, 1, pack=pks, index=torch.tensor([3, 0, 2, 1], dtype=torch.int32, device=dev), want_status=True)
A = torch.randn(3, 160, 200, device=dev); A[1, 150:] = 0; c = torch.randn(3, 200, device=dev)
out = cave_forward_backward(c, A, -1.0, 0, want_proj=True, want_status=True, dense_slots=2)          # dense path, two rounds
assert (out["status"] & 0x200).all()
sol, obj, tours = tsp_exact.solve(torch.rand(3, 28) + 0.1, 8)
torch.cuda.synchronize()
print("sanitize case done", float(out["loss"]), obj)
