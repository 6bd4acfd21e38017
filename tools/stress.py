"""Stress campaign: many seeds x workloads x prediction regimes; every instance must converge (status 0 / skipped)
and satisfy the projection's optimality conditions, checked on the device in float64:
    q = c - p in the polar cone (A q <= tol),  <p, q> = 0,  rnorm = ||q||."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from cave_b200 import cave_forward_backward, synth

dev = torch.device("cuda:0")
rng = np.random.default_rng(2024)
tot = bad = 0
worst = dict(polar=0.0, orth=0.0)
t0 = time.time()
cases = [("sp5", 2048), ("tsp20", 1024), ("vrp20", 1024), ("tsp50", 296), ("tsp10", 2048), ("tsp35", 512),
         ("dense15x10", 1024), ("dense40x12", 1024), ("dense64x190", 296), ("dense30x30", 1024)]
regimes = ["uniform", "near", "gauss", "scaled", "sparse"]
for rep in range(int(sys.argv[1]) if len(sys.argv) > 1 else 3):
    for kind, B in cases:
        seed = int(rng.integers(1 << 30))
        if kind.startswith("dense"):
            m, d = (int(x) for x in kind[5:].split("x"))
            A = torch.tensor(np.random.default_rng(seed).standard_normal((B, m, d)).astype(np.float32), device=dev)
            A[:, m - 2:, :] = 0.0                                   # padding rows
            insts = None
        else:
            insts = synth.make_batch(kind, B, seed=seed)
            A = synth.densify(insts, device=dev)
        for regime in regimes:
            if insts is None and regime in ("uniform", "near"):
                continue
            if regime in ("uniform", "near"):
                pred = torch.tensor(synth.predictions(insts, seed, regime), device=dev, dtype=torch.float64)
            else:
                r2 = np.random.default_rng(seed + 1)
                pnp = r2.standard_normal((B, A.shape[2]))
                if regime == "scaled":
                    pnp = pnp * np.logspace(-6, 6, B)[:, None]
                if regime == "sparse":
                    pnp = pnp * (r2.random(pnp.shape) < 0.1)
                pred = torch.tensor(pnp, device=dev, dtype=torch.float64)
            for prec in ("fp64", "fp32"):
                out = cave_forward_backward(pred, A, -1.0, 1, 0.2, "none", precision=prec, want_proj=True, want_status=True)
                st = out["status"] & 0xff
                ok = (st == 0) | (st == 4)
                c, p = -pred, out["proj"]
                q = c - p
                cn = c.norm(dim=1).clamp(min=1e-300)
                Aq = torch.bmm(A.double(), q.unsqueeze(2)).squeeze(2).max(dim=1).values / cn
                orth = (p * q).sum(1).abs() / (cn * cn)
                viol = (Aq > 1e-7) | (orth > 1e-7) | ~torch.isfinite(out["loss"]) | ~ok
                tot += B
                nb = int(viol.sum())
                bad += nb
                worst["polar"] = max(worst["polar"], float(Aq.max())); worst["orth"] = max(worst["orth"], float(orth.max()))
                if nb:
                    i = int(torch.nonzero(viol)[0])
                    print(f"VIOLATION {kind} seed {seed} {regime} {prec}: {nb} instances, e.g. #{i} status {int(st[i])} iters "
                          f"{int(out['iters'][i])} polar {float(Aq[i]):.2e} orth {float(orth[i]):.2e}", flush=True)
print(f"{tot} instance-solves, {bad} violations, worst polar {worst['polar']:.2e} worst orth {worst['orth']:.2e}, {time.time()-t0:.0f} s")
