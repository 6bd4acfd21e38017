"""Probe: does splitting a batch into independent parts on separate streams (scan of one part overlapping the solve
of another) beat one stream?  TSP-50, B = 4096, fp64, cold pack.  Usage: python tools/overlap_probe.py [parts ...]"""
import sys

import torch

from cave_b200 import cave_forward_backward, synth

dev = torch.device("cuda:0")
B = 4096
insts = synth.make_batch("tsp50", B, seed=0)
A = synth.densify(insts, device=dev)
pred = torch.as_tensor(synth.predictions(insts, 0, "uniform"), dtype=torch.float32, device=dev)


def step(parts, streams):
    n = B // parts
    ev = torch.cuda.Event(); ev.record()
    outs = []
    for i in range(parts):
        s = streams[i % len(streams)]
        s.wait_event(ev)
        with torch.cuda.stream(s):
            outs.append(cave_forward_backward(pred[i * n:(i + 1) * n], A[i * n:(i + 1) * n], -1.0, 1, 0.2, "sum", precision="fp64"))
    for s in streams:
        torch.cuda.current_stream().wait_stream(s)
    return outs


for parts, nstreams in [(1, 1), (2, 2), (4, 2), (4, 4), (8, 2), (8, 4), (2, 1), (4, 1)]:
    streams = [torch.cuda.Stream() for _ in range(nstreams)]
    for _ in range(3):
        step(parts, streams)
    torch.cuda.synchronize()
    t0, t1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0.record()
    for _ in range(5):
        step(parts, streams)
    t1.record(); torch.cuda.synchronize()
    ms = t0.elapsed_time(t1) / 5
    print(f"parts {parts} streams {nstreams}: {ms:.3f} ms/step  {B / ms * 1e3:.0f} inst/s")
