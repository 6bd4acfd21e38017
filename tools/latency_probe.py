"""Per-call latency of the module at the reference's small batch sizes (batch 32)."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from cave_b200 import EPO, innerConeAlignedCosine, cave_forward_backward, pack_constraints, synth

class M: modelSense = EPO.MINIMIZE
dev = torch.device("cuda:0")
for kind in ("sp5", "tsp20", "tsp50"):
    B = 32
    insts = synth.make_batch(kind, B, seed=5)
    A = synth.densify(insts, device=dev)
    pred = torch.tensor(synth.predictions(insts, 5, "near"), device=dev, requires_grad=True)
    mod = innerConeAlignedCosine(M(), solver="cuda", seed=0)
    pack = pack_constraints(A)
    idx = torch.arange(B, dtype=torch.int32, device=dev)
    def dense():
        loss = mod(pred, A); loss.backward(); return loss
    def resident():
        loss = mod(pred, pack, index=idx); loss.backward(); return loss
    def raw():
        return cave_forward_backward(pred.detach(), A, -1.0, 1, 0.2, "mean")
    for name, fn in (("module dense", dense), ("module resident", resident), ("raw call", raw)):
        for _ in range(5): fn()
        torch.cuda.synchronize(); t0 = time.perf_counter()
        n = 50
        for _ in range(n): fn()
        torch.cuda.synchronize(); dt = (time.perf_counter() - t0) / n
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(n): fn()
        e1.record(); torch.cuda.synchronize()
        print(f"{kind:6s} B={B} {name:16s}: wall {dt*1e6:8.1f} us/call   device {e0.elapsed_time(e1)/n*1e3:8.1f} us/call")
