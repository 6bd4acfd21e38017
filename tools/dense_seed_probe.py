"""Which instances of a seeded dense batch leave the Gram path (status / iterations), for robustness work."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from cave_b200 import cave_forward_backward
dev = torch.device("cuda:0")
d, m, B = (int(x) for x in sys.argv[1:4])
for seed in [int(s) for s in sys.argv[4:]] or [1000, 1001]:
    g = torch.Generator(device=dev).manual_seed(seed)
    A = torch.randn((B, m, d), generator=g, device=dev)
    c = -torch.randn((B, d), generator=g, device=dev)
    torch.cuda.synchronize(); t0 = time.perf_counter()
    out = cave_forward_backward(c, A, -1.0, 0, 0.0, "none", want_status=True)
    torch.cuda.synchronize(); dt = time.perf_counter() - t0
    st = out["status"].cpu().numpy(); it = out["iters"].cpu().numpy()
    vals, cnts = np.unique(st, return_counts=True)
    print(f"seed {seed}: {B / dt:.1f} inst/s; status counts {dict(zip([hex(v) for v in vals], cnts.tolist()))}; iters mean {it.mean():.1f} max {it.max()}", flush=True)
    bad = np.flatnonzero(((st & 0x200) == 0) | ((st >> 12) != 0))
    print("   instances off the Gram path / flagged:", bad[:20].tolist(), "iters", it[bad][:20].tolist(), "status", [hex(x) for x in st[bad][:20]])
    if len(bad) and os.environ.get("DUMP"):
        i = int(bad[0])
        np.savez_compressed(f"gpurun_out/dense_bad_{seed}_{i}.npz", A=A[i].cpu().numpy(), c=c[i].cpu().numpy(),
                            proj=cave_forward_backward(c[i:i+1], A[i:i+1], -1.0, 0, 0.0, "none", want_proj=True, dense=False)["proj"].cpu().numpy())
