"""Dense sweep shapes (BASELINE.json configs[4]): correctness vs scipy on a few instances + GPU timing."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from cave_b200 import cave_forward_backward
from oracle import cave_oracle as O

dev = torch.device("cuda:0")
shapes = [(190, 64), (190, 256), (190, 1024), (190, 2048), (1225, 64), (1225, 256), (1225, 1024), (4950, 64), (4950, 256)]
if len(sys.argv) > 1:
    shapes = [tuple(int(x) for x in a.split("x")) for a in sys.argv[1:]]
for d, m in shapes:
    B = 296
    g = torch.Generator(device=dev).manual_seed(d * 7 + m)
    A = torch.randn((B, m, d), generator=g, device=dev)
    c = torch.randn((B, d), generator=g, device=dev, dtype=torch.float64)
    for prec in ("fp64",):
        cave_forward_backward(c, A, 1.0, 0, precision=prec)
        torch.cuda.synchronize(); t0 = time.perf_counter()
        out = cave_forward_backward(c, A, 1.0, 0, reduction="none", precision=prec, want_proj=True, want_status=True)
        torch.cuda.synchronize(); dt = time.perf_counter() - t0
        st = out["status"].cpu().numpy(); it = out["iters"].cpu().numpy()
        n = 2
        ref_p, ref_r = O.batch_project(c[:n].cpu().numpy(), A[:n].cpu().numpy(), fp64=True)
        err = np.abs(out["proj"][:n].cpu().numpy() - ref_p).max() / max(np.abs(ref_p).max(), 1e-30)
        print(f"d={d:5d} m={m:5d} {prec}: {B/dt:10.1f} inst/s  ({dt*1e3:8.1f} ms for {B})  status {sorted(set(st.tolist()))} "
              f"iters mean {it.mean():.0f} max {it.max()}  relerr vs scipy {err:.1e}", flush=True)
