"""SP5 step time on every rank before / after the NCCL communicator exists (torchrun, one rank per GPU)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch, torch.distributed as dist
from cave_b200 import cave_forward_backward, synth
rank, local = int(os.environ["RANK"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
kind, B, mode = sys.argv[1], int(sys.argv[2]), int(sys.argv[3])
insts = synth.make_batch(kind, B, seed=1000 + rank)
A = synth.densify(insts, device=dev)
pred = torch.tensor(synth.predictions(insts, 1000 + rank, "uniform"), device=dev)

def t(tag):
    for _ in range(3):
        cave_forward_backward(pred, A, -1.0, mode, 0.2, "mean")
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize(); e0.record()
    for _ in range(10):
        cave_forward_backward(pred, A, -1.0, mode, 0.2, "mean")
    e1.record(); torch.cuda.synchronize()
    print(f"rank {rank} {tag}: {e0.elapsed_time(e1) / 10:.3f} ms/step", flush=True)

t("before init_process_group")
dist.init_process_group("nccl", device_id=dev) if os.environ.get("EAGER") else dist.init_process_group("nccl")
t("after init_process_group (no collective yet)")
x = torch.ones(1, device=dev); dist.all_reduce(x); torch.cuda.synchronize()
t("after the first all_reduce")
dist.barrier(); torch.cuda.synchronize()
t("after a barrier")
dist.destroy_process_group()
t("after destroy_process_group")
