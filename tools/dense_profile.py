"""Coarse phase breakdown of the dense solve kernel (clock64 deltas of thread 0, summed over instances).
Needs the profile build:  CAVE_NVCC_EXTRA=-DCAVE_DENSE_PROFILE python cave_b200/build.py libcave_b200_prof.so
and  CAVE_B200_LIB=cave_b200/_C/libcave_b200_prof.so CAVE_KEEP_SCRATCH=1 python tools/dense_profile.py d m B"""
import ctypes, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from cave_b200 import _lib, cave_forward_backward

NAMES = ["setup", "kkt", "freeset", "gather", "chol", "trisolve", "ls_gram", "ls_true", "truegrad", "epilogue", "switch", "n_fact", "n_iter",
         "chol:update", "chol:diag", "chol:trsm"]
dev = torch.device("cuda:0")
d, m, B = (int(x) for x in sys.argv[1:4])
g = torch.Generator(device=dev).manual_seed(d * 7 + m)
A = torch.randn((B, m, d), generator=g, device=dev)
c = torch.randn((B, d), generator=g, device=dev, dtype=torch.float64)
cave_forward_backward(c, A, 1.0, 0, reduction="none", dense=True)
torch.cuda.synchronize(); t0 = time.perf_counter()
out = cave_forward_backward(c, A, 1.0, 0, reduction="none", want_status=True, dense=True)
torch.cuda.synchronize(); dt = time.perf_counter() - t0
off = ctypes.c_size_t()
_lib.check(_lib.load().cave_dense_ctrl_offset(B, m, d, ctypes.byref(out["_opts"]), ctypes.byref(off)))
blk = out["_scratch"][off.value:off.value + 192].cpu()
ints = blk[:64].view(torch.int32).tolist(); prof = blk[64:192].view(torch.int64).tolist()
tot = sum(prof[:11])
print(f"d={d} m={m} B={B}: {B / dt:.1f} inst/s ({dt * 1e3:.1f} ms), dense instances {ints[0]}, iters mean {out['iters'].float().mean():.1f}")
for n, v in zip(NAMES, prof):
    if n.startswith("chol:"):
        print(f"  {n:12s} {v / max(ints[0], 1) / 1e3:10.1f} kclk per instance")
    elif n.startswith("n_"):
        print(f"  {n:10s} {v / max(ints[0], 1):8.2f} per instance")
    else:
        print(f"  {n:10s} {v / max(ints[0], 1) / 1e3:10.1f} kclk per instance  {100.0 * v / max(tot, 1):5.1f} %")
