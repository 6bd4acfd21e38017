"""Dense regime bring-up: Gram kernel vs a float64 product, then the dense path vs the oracle, with timings.
   python tools/dense_check.py [gram|solve|time] [d m B]..."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from cave_b200 import cave_forward_backward
from cave_b200.qpsolver import dense_gram
from oracle import cave_oracle as O

dev = torch.device("cuda:0")
what = sys.argv[1] if len(sys.argv) > 1 else "gram"
args = [int(x) for x in sys.argv[2:]]
shapes = [tuple(args[i:i + 3]) for i in range(0, len(args), 3)] or [(64, 128, 2), (190, 256, 3), (333, 300, 2), (1225, 1024, 2)]
for d, m, B in shapes:
    g = torch.Generator(device=dev).manual_seed(d * 7 + m)
    A = torch.randn((B, m, d), generator=g, device=dev)
    if what == "gram":
        if B > 1 and m > 140:
            A[1, m - 9:] = 0          # ragged: fewer valid rows in instance 1
        G, cnt = dense_gram(A)
        torch.cuda.synchronize()
        for b in range(B):
            mv = int((A[b].abs().sum(1) > 0).sum())
            ref = A[b, :mv].double() @ A[b, :mv].double().T
            got = G[b, :mv, :mv].double()
            nrm = A[b, :mv].double().norm(dim=1)
            err = ((got - ref).abs() / (nrm[:, None] * nrm[None, :])).max().item()
            asym = (G[b, :mv, :mv] - G[b, :mv, :mv].T).abs().max().item()
            print(f"gram d={d} m={m} inst {b} (valid {mv}): dense count {cnt}, max err / (|a_i||a_j|) = {err:.2e}, asym {asym:.1e}", flush=True)
    else:
        c = torch.randn((B, d), generator=g, device=dev, dtype=torch.float64)
        for dense in ((True,) if what == "solve" else (True, False)):
            out = cave_forward_backward(c, A, 1.0, 0, reduction="none", want_proj=True, want_status=True, dense=dense)
            torch.cuda.synchronize(); t0 = time.perf_counter()
            out = cave_forward_backward(c, A, 1.0, 0, reduction="none", want_proj=True, want_status=True, dense=dense)
            torch.cuda.synchronize(); dt = time.perf_counter() - t0
            st = out["status"].cpu().numpy(); it = out["iters"].cpu().numpy()
            line = f"d={d} m={m} B={B} dense={dense}: {B / dt:9.1f} inst/s ({dt * 1e3:.1f} ms) status {sorted(set(hex(x) for x in st.tolist()))} iters mean {it.mean():.1f} max {it.max()}"
            if what == "solve":
                n = min(B, 2)
                ref_p, ref_r = O.batch_project(c[:n].cpu().numpy(), A[:n].cpu().numpy(), fp64=True)
                err = np.abs(out["proj"][:n].cpu().numpy() - ref_p).max() / max(np.abs(ref_p).max(), 1e-30)
                line += f" relerr vs scipy {err:.1e} rnorm err {np.abs(out['rnorm'][:n].cpu().numpy() - ref_r).max():.1e}"
            print(line, flush=True)
