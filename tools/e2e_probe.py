"""Where the time of one device-resident end-to-end step goes (host side): pinned pred in, loss + gradient out."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from cave_b200 import EPO, innerConeAlignedCosine, pack_constraints, cave_forward_backward, synth

dev = torch.device("cuda:0")
B = 4096
insts = synth.make_batch("tsp50", B, seed=1000)
A = synth.densify(insts, device=dev)
pred = torch.tensor(synth.predictions(insts, 1000, "uniform"), device=dev)
pack = pack_constraints(A)
class M: modelSense = EPO.MINIMIZE
mod = innerConeAlignedCosine(M(), solver="cuda", inner_ratio=0.2, seed=0)
pred_host = torch.empty(pred.shape, dtype=pred.dtype, pin_memory=True); pred_host.copy_(pred)
perm_host = torch.randperm(B, dtype=torch.int32).pin_memory()
perm_dev = perm_host.to(dev)

def t(fn, n=20):
    fn(); torch.cuda.synchronize(); t0 = time.perf_counter()
    for _ in range(n): fn()
    torch.cuda.synchronize(); return (time.perf_counter() - t0) / n * 1e3

def full():
    p = pred_host.requires_grad_(True); p.grad = None
    loss = mod(p, pack, index=perm_host); loss.backward(); return loss.item()
def dev_only():
    return cave_forward_backward(pred, None, -1.0, 1, 0.2, "mean", pack=pack, index=perm_dev)
def dev_module():
    p = pred.detach().requires_grad_(True)
    loss = mod(p, pack, index=perm_dev); loss.backward(); return loss
def h2d():
    pred_host.to(dev, non_blocking=True); perm_host.to(dev, non_blocking=True)
g = torch.empty_like(pred)
def d2h():
    gh = torch.empty(g.shape, dtype=g.dtype, pin_memory=True); gh.copy_(g, non_blocking=True); torch.cuda.current_stream().synchronize()
print(f"full host->host step          {t(full):7.3f} ms")
print(f"C-ABI call, device tensors    {t(dev_only):7.3f} ms")
print(f"module fwd+bwd, device tensors{t(dev_module):7.3f} ms")
print(f"H2D pred + index (pinned)     {t(h2d):7.3f} ms")
print(f"D2H gradient into fresh pinned{t(d2h):7.3f} ms")
import cProfile, pstats
pr = cProfile.Profile(); pr.enable()
for _ in range(20): full()
pr.disable(); pstats.Stats(pr).sort_stats("cumulative").print_stats(18)
