"""One forward+backward of the OFF engine at the bench shape, for ncu (python tools/prof_step.py [B] [L] [prec])."""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import off_b200  # noqa
from off_b200 import engine as E

B = int(sys.argv[1]) if len(sys.argv) > 1 else 48
Lg = int(sys.argv[2]) if len(sys.argv) > 2 else 3
prec = sys.argv[3] if len(sys.argv) > 3 else "tf32"
iters = int(sys.argv[4]) if len(sys.argv) > 4 else 1
eng = E.OFFEngine(B, Lg, "rgb", "cuda", prec)
torch.manual_seed(0)
with torch.no_grad():
    for n, v in eng.params.items():
        fan = v[0].numel() if v.dim() > 1 else 64
        v.uniform_(-1.0 / fan ** 0.5, 1.0 / fan ** 0.5)
for t in eng.taps.values():
    t.copy_(torch.relu(torch.randn_like(t)))
g7 = torch.randn(eng.P, 101, device="cuda") * 0.01
g14 = torch.randn(eng.P, 101, device="cuda") * 0.01
import time
eng.single_stream = os.environ.get("OFFK_SINGLE_STREAM", "0") == "1"
for i in range(iters):
    if i == iters - 1:                      # ncu --profile-from-start off: only the last iteration is profiled
        torch.cuda.synchronize()
        torch.cuda.cudart().cudaProfilerStart()
    t0 = time.perf_counter()
    eng.forward(train=True, seed=1)
    eng.backward(g7, g14)
    t1 = time.perf_counter()
    torch.cuda.synchronize()
    t2 = time.perf_counter()
    print(f"iter {i}: CPU issue {1e3*(t1-t0):.2f} ms, until done {1e3*(t2-t0):.2f} ms")
torch.cuda.cudart().cudaProfilerStop()
with open(os.path.join(ROOT, "gpurun_out", "step_names.txt"), "w") as f:
    f.write("\n".join(["seed_set"] + eng.launch_names()))      # forward(train=True) first stores the step seed (one tiny kernel)
print("done")
