"""Per-phase cycle breakdown of the solve kernel (CAVE_PROFILE=1 debug counters)."""
import ctypes, os, sys
os.environ["CAVE_PROFILE"] = "1"
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from cave_b200 import _lib, cave_forward_backward, pack_constraints, synth

NAMES = ["0 load c", "1 csr load", "2 merge", "3 vars", "4 csc", "5 eval0", "6 grad", "7 res/freelist", "8 hessian",
         "9 copy L", "10 ldlt", "11 backsolve", "12 dir+linesearch", "13 exit", "14 epilogue", "15 fetch/idle"]
kind = sys.argv[1] if len(sys.argv) > 1 else "tsp50"
B = int(sys.argv[2]) if len(sys.argv) > 2 else 1184
prec = sys.argv[3] if len(sys.argv) > 3 else "fp32"
dev = torch.device("cuda:0")
insts = synth.make_batch(kind, B, seed=1000)
A = synth.densify(insts, device=dev)
pred = torch.tensor(synth.predictions(insts, 1000, "uniform"), device=dev)
pack = pack_constraints(A)
lib = _lib.load()
buf = (ctypes.c_ulonglong * 32)()
for _ in range(2):
    cave_forward_backward(pred, A, -1.0, 1, 0.2, "mean", precision=prec, pack=pack)
lib.cave_debug_phase_cycles(buf, 1)
out = cave_forward_backward(pred, A, -1.0, 1, 0.2, "mean", precision=prec, pack=pack, want_status=True)
lib.cave_debug_phase_cycles(buf, 1)
cyc = np.array(list(buf), dtype=np.float64)
it = out["iters"].float().mean().item()
tot = cyc.sum()
print(f"{kind} B={B} {prec}: mean iters {it:.2f}; total {tot / B:.0f} cycles/instance")
for n, c in zip(NAMES, cyc):
    print(f"  {n:20s} {c / B:10.0f} cyc/inst  {100 * c / tot:5.1f}%")
