import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from cave_b200 import cave_forward_backward
z = np.load(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests/golden/regress/sp5_newton_floor.npz"))
A = torch.tensor(z["A"][None], device="cuda:0"); pred = torch.tensor(z["pred"][None], device="cuda:0")
out = cave_forward_backward(pred, A, -1.0, 0, 0.2, "none", want_status=True)
torch.cuda.synchronize()
print("status", hex(int(out["status"][0])), "iters", int(out["iters"][0]))
