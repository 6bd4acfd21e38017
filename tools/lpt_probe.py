"""How much would ordering the solve kernel's work queue by a cost estimate help?  Simulates the persistent-CTA
queue (444 workers) with a per-instance time model fitted to the measured kernel time."""
import heapq
import numpy as np
import torch
from cave_b200 import cave_forward_backward, pack_constraints, synth

dev = torch.device("cuda:0")
B = 4096
insts = synth.make_batch("tsp50", B, seed=0)
A = synth.densify(insts, device=dev)
pred = torch.as_tensor(synth.predictions(insts, 0, "uniform"), dtype=torch.float32, device=dev)
out = cave_forward_backward(pred, A, -1.0, 1, 0.2, "mean", precision="fp64", want_status=True)
iters = out["iters"].cpu().numpy().astype(float)
nnz = np.array([len(i.vals) - i.d for i in insts], dtype=float)       # non-zeros of the general rows
t = 60.0 + iters * (20.0 + 0.006 * nnz) + 0.004 * nnz                  # microseconds, shape of the cost only


def makespan(order, workers=444):
    h = [0.0] * workers
    heapq.heapify(h)
    for i in order:
        heapq.heappush(h, heapq.heappop(h) + t[i])
    return max(h)


ideal = t.sum() / 444
print("ideal %.1f  fifo %.1f (+%.1f%%)  by nnz %.1f (+%.1f%%)  oracle LPT %.1f (+%.1f%%)" % (
    ideal, makespan(range(B)), 100 * (makespan(range(B)) / ideal - 1),
    makespan(np.argsort(-nnz)), 100 * (makespan(np.argsort(-nnz)) / ideal - 1),
    makespan(np.argsort(-t)), 100 * (makespan(np.argsort(-t)) / ideal - 1)))
print("iters mean %.2f std %.2f; nnz mean %.0f std %.0f; corr(t, nnz) %.2f" % (iters.mean(), iters.std(), nnz.mean(), nnz.std(), np.corrcoef(t, nnz)[0, 1]))
