"""Per-CTA phase timeline of the TMA-fed GEMM launches of one step (diagnostics build only):

    OFFK_OUT=$PWD/gpurun_out/liboffk_tl.so OFFK_OBJDIR=/tmp/offk_tl OFFK_EXTRA_FLAGS=-DOFFK_TIMELINE bash .../csrc/build.sh
    OFFK_LIB=$PWD/gpurun_out/liboffk_tl.so python tools/timeline.py 48 3 fp32 [name-substring ...]

For every matching launch: the median / p90 over its CTAs of the clocks spent in each phase
(entry -> prologue -> dependency wait -> first K-block ready -> last MMA issued -> accumulator complete -> epilogue done).
Says where a short GEMM's time goes; nothing here is a bench number (single stream, sync after every launch).
"""
import ctypes as C
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import off_b200  # noqa
from off_b200 import engine as E
from off_b200 import _lib as L

B = int(sys.argv[1]) if len(sys.argv) > 1 else 48
Lg = int(sys.argv[2]) if len(sys.argv) > 2 else 3
prec = sys.argv[3] if len(sys.argv) > 3 else "fp32"
pats = sys.argv[4:]
eng = E.OFFEngine(B, Lg, "rgb", "cuda", prec)
lib = eng.lib
torch.manual_seed(0)
with torch.no_grad():
    for n, v in eng.params.items():
        fan = v[0].numel() if v.dim() > 1 else 64
        v.uniform_(-1.0 / fan ** 0.5, 1.0 / fan ** 0.5)
for t in eng.taps.values():
    t.copy_(torch.relu(torch.randn_like(t)))
g7 = torch.randn(eng.P, 101, device="cuda") * 0.01
g14 = torch.randn(eng.P, 101, device="cuda") * 0.01
eng.single_stream = True
for _ in range(2):
    eng.forward(train=True, seed=1)
    eng.backward(g7, g14)
torch.cuda.synchronize()

raw = C.CDLL(L.LIB_PATH)
raw.offk_timeline_read.argtypes = [C.c_void_p, C.c_int]
SLOTS, MAXC = 32, 8192
host = np.zeros(MAXC * SLOTS, dtype=np.int64)
PH = ["prologue", "dep wait", "first kb", "main loop", "split tail", "acc wait", "epilogue"]
print(f"# B={B} L={Lg} {prec}; clocks per phase, median (p90) over CTAs; 'span' = entry -> epilogue done")
print(f"# {'launch':36s} {'ctas':>5s} " + " ".join(f"{p:>14s}" for p in ["prologue", "dep wait", "->first kb", "->last mma", "->acc done", "epilogue", "span", "c1:tmem", "c1:sts", "c1:bar", "c1:ld+st"]))


def on_step(step, stream):
    name = getattr(step, "name", None) or "/".join(getattr(step, "launches", ["?"]))
    torch.cuda.synchronize()
    raw.offk_timeline_read(host.ctypes.data, MAXC)
    a = host.reshape(MAXC, SLOTS)
    live = a[:, 7] > 0
    if not live.any():
        return
    raw.offk_timeline_clear()
    if pats and not any(p in name for p in pats):
        return
    a = a[live]
    t0, t1, t2, t3, t4, t5, t6, t7 = [a[:, i].astype(np.float64) for i in range(8)]
    cols = [t1 - t0, t2 - t1, t3 - t2, t4 - t3, t6 - t4, t7 - t6, t7 - t0]
    if (a[:, 12] > 0).all():        # chunk 1 of the vector epilogue: TMEM drain, staging, barrier, global loads + stores
        c8, c9, c10, c11, c12 = [a[:, i].astype(np.float64) for i in range(8, 13)]
        cols += [c8 - c11, c9 - c8, c10 - c9, c12 - c10]
    fmt = lambda x: f"{np.median(x):7.0f} ({np.percentile(x, 90):6.0f})"
    print(f"{name:38s} {a.shape[0]:5d} " + " ".join(f"{fmt(c):>14s}" for c in cols))
    if (a[:, 25] > 0).all():        # steady state (K-blocks 8 and 9): one stage's trip round the pipeline
        q = lambda i: a[:, i].astype(np.float64)
        st = {"empty->tma issued": q(17) - q(16), "tma issued->landed": q(18) - q(17), "split": q(19) - q(18),
              "split->mma warp": q(20) - q(19), "mma issue": q(21) - q(20), "period (mma kb8->kb9)": q(25) - q(21),
              "producer period": q(22) - q(16), "split period": q(23) - q(18), "first split": q(14) - q(13)}
        print("      steady: " + "  ".join(f"{k} {np.median(v):.0f}" for k, v in st.items()))


raw.offk_timeline_clear()
torch.cuda.synchronize()
eng._set_dropout(True, None, 1)
streams = eng._fork()
eng.fwd_sched.run(streams, on_step=on_step)
eng.d_out7.copy_(g7.reshape(eng.d_out7.shape))
eng.d_out14.copy_(g14.reshape(eng.d_out14.shape))
eng._zero_grads = True
eng.bwd_sched.run(streams, on_step=on_step)
torch.cuda.synchronize()
print("done")
