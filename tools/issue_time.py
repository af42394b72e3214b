"""Host-issue vs device time of one OFF forward+backward at the bench shape, without a profiler
(python tools/issue_time.py [B] [L] [prec]).  Says whether a step is bound by the Python/ctypes launch loop."""
import os
import sys
import time

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import off_b200  # noqa
from off_b200 import engine as E

B = int(sys.argv[1]) if len(sys.argv) > 1 else 48
Lg = int(sys.argv[2]) if len(sys.argv) > 2 else 3
prec = sys.argv[3] if len(sys.argv) > 3 else "tf32"
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


def run(n, what, graph=False):
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0 = time.perf_counter()
    a.record()
    for i in range(n):
        if what in ("both", "fwd"):
            eng.forward(train=True, seed=i + 1, graph=graph)
        if what in ("both", "bwd"):
            eng.backward(g7, g14, graph=graph)
    b.record()
    t1 = time.perf_counter()
    torch.cuda.synchronize()
    return 1e3 * (t1 - t0) / n, a.elapsed_time(b) / n


for single in (False, True):
    eng.single_stream = single
    run(3, "both")
    for what in ("fwd", "bwd", "both"):
        cpu, gpu = run(10, what)
        print(f"B={B} L={Lg} {prec} single_stream={single} {what:4s}: host issue {cpu:7.3f} ms/step, device {gpu:7.3f} ms/step "
              f"({eng.launches_fwd}+{eng.launches_bwd} launches)")

eng.single_stream = False
run(3, "both", True)                      # captures
for what in ("fwd", "bwd", "both"):
    cpu, gpu = run(10, what, True)
    print(f"B={B} L={Lg} {prec} CUDA GRAPH replay {what:4s}: host issue {cpu:7.3f} ms/step, device {gpu:7.3f} ms/step")
