"""Status / iteration histogram and step time of one structured workload for several generator seeds (a slow seed points at a
data-dependent slow path).  usage: python tools/seed_probe_structured.py kind B mode seed [seed ...]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from cave_b200 import cave_forward_backward, synth
dev = torch.device("cuda:0")
kind, B, mode = sys.argv[1], int(sys.argv[2]), int(sys.argv[3])
for seed in [int(s) for s in sys.argv[4:]]:
    insts = synth.make_batch(kind, B, seed=seed)
    A = synth.densify(insts, device=dev)
    pred = torch.tensor(synth.predictions(insts, seed, os.environ.get("REGIME", "near")), device=dev)
    for _ in range(2):
        out = cave_forward_backward(pred, A, -1.0, mode, 0.2, "mean", want_status=True)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(5):
        out = cave_forward_backward(pred, A, -1.0, mode, 0.2, "mean", want_status=True)
    e1.record(); torch.cuda.synchronize()
    st = out["status"].cpu().numpy(); it = out["iters"].cpu().numpy()
    vals, cnts = np.unique(st, return_counts=True)
    print(f"{kind} seed {seed}: shape {tuple(A.shape)} {e0.elapsed_time(e1) / 5:.3f} ms/step; status {dict(zip([hex(v) for v in vals], cnts.tolist()))}; "
          f"iters mean {it.mean():.1f} max {it.max()}; slowest instances {np.argsort(-it)[:5].tolist()}", flush=True)
