"""A/B timing of dense-path knobs inside ONE process (same box, same clocks): alternates environment settings that the
library reads per launch (e.g. CAVE_DENSE_TC=1 / 0) and prints the median device time of each.
usage: python tools/dense_ab.py d m B reps VAR=a,b [VAR2=c,d ...]   (settings are zipped, not crossed; DENSE_SLOTS=n is passed as
dense_slots)"""
import os, sys, statistics
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from cave_b200 import cave_forward_backward

dev = torch.device("cuda:0")
d, m, B, reps = (int(x) for x in sys.argv[1:5])
knobs = [a.split("=") for a in sys.argv[5:]]
names = [k for k, _ in knobs]
settings = list(zip(*[v.split(",") for _, v in knobs])) or [()]
g = torch.Generator(device=dev).manual_seed(d * 7 + m)
A = torch.randn((B, m, d), generator=g, device=dev)
c = torch.randn((B, d), generator=g, device=dev, dtype=torch.float64)
times = {s: [] for s in settings}
ref = None
for rep in range(reps + 1):
    for s in settings:
        for k, v in zip(names, s): os.environ[k] = v
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        out = cave_forward_backward(c, A, 1.0, 0, reduction="none", want_status=True, dense=True,
                                    dense_slots=int(os.environ.get("DENSE_SLOTS", "0")) or None)
        e1.record(); torch.cuda.synchronize()
        if rep: times[s].append(e0.elapsed_time(e1))
        if ref is None: ref = out["grad"].clone()
        err = float((out["grad"] - ref).abs().max())
        assert err < 1e-6, (s, err)
for s in settings:
    t = statistics.median(times[s])
    print(dict(zip(names, s)), f"median {t:.2f} ms -> {B / t * 1e3:.0f} inst/s   (min {min(times[s]):.2f}, max {max(times[s]):.2f})")
