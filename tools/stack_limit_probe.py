"""Does the device's stack-size limit (NCCL raises it when a communicator is created) change the solve kernel's speed?
usage: python tools/stack_limit_probe.py BYTES kind B mode seed..."""
import ctypes, sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
torch.zeros(1, device="cuda:0")
rt = ctypes.CDLL("libcudart.so.12")
cur = ctypes.c_size_t()
rt.cudaDeviceGetLimit(ctypes.byref(cur), 0)          # cudaLimitStackSize = 0
print("stack limit before:", cur.value)
want = int(sys.argv[1])
if want > 0:
    print("cudaDeviceSetLimit ->", rt.cudaDeviceSetLimit(0, ctypes.c_size_t(want)))
rt.cudaDeviceGetLimit(ctypes.byref(cur), 0)
print("stack limit now:", cur.value)
sys.argv = ["x"] + sys.argv[2:]
exec(open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "seed_probe_structured.py")).read())
