"""OFF stencil / temporal-difference microbenchmark (BASELINE.json config 5).

    python tools/bench_stencil.py [--sweep] [--batch 48] [--length 3] [--reps 10] [--json out.json]

Times offk_stencil_diff_fwd_batch / _bwd_batch through the C ABI with CUDA events on the launching stream, L2
evicted before every launch (write 256 MB, then read 256 MB so the dirty lines are written back before the timed
launch).  Reports algorithmic GB/s = 4*S^2*(Cg*N + Cs*P + (K*Cs+Cg)*P) / time  (forward) and
4*S^2*((K*Cs+Cg)*P + Cg*N + Cg*N + Cs*P) / time (backward), against MEASURED_PEAKS.json's HBM copy bandwidth.
Default: the three per-stage launches of the reference shapes (config 2).  --sweep: S in {28,14,7}, Cg in
{128,256,512,1024}, with B*L grown until the working set is >= 1 GB (far beyond the 126 MB L2).
"""
import argparse
import ctypes as C
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import off_b200  # noqa: E402,F401
from off_b200 import _lib as L  # noqa: E402


def peak():
    try:
        return json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"], "measured"
    except Exception:
        return 6650.0, "fallback"


class Flusher:
    def __init__(self, dev):
        self.w = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
        self.r = torch.zeros(256 << 20, dtype=torch.uint8, device=dev)
        self.sink = torch.zeros((), dtype=torch.int64, device=dev)

    def __call__(self):
        self.w.fill_(1)
        self.sink.add_(self.r.view(torch.int64).sum())


def make_level(dev, B, Lg, S, Cg, Cs, ctot, coff, drop, learned=True):
    N, P = B * Lg, B * (Lg - 1)
    cu = Cg + Cs
    gd = torch.randn(N, S, S, cu, device=dev)
    gd[..., :Cg].relu_()
    dgd = torch.empty_like(gd)
    w = torch.randn(Cs, 1, 3, 3, device=dev)
    bias = torch.randn(Cs, device=dev)
    dw, db = torch.zeros_like(w), torch.zeros_like(bias)
    sd = L.OffkStencil()
    sd.B, sd.L, sd.Cg, sd.Cs, sd.K, sd.H, sd.W = B, Lg, Cg, Cs, 1, S, S
    sd.g_fs = sd.d_fs = cu * S * S
    sd.g_ps = sd.d_ps = cu
    sd.out_ctot, sd.out_coff, sd.index_mode = ctot, coff, L.INDEX_REFERENCE_FLAT
    sd.drop_mode, sd.keep_scale, sd.drop_p, sd.seed = drop, 5.0, 0.8, 99
    keep = (gd, dgd, w, bias, dw, db)
    fb = 4.0 * S * S * (Cg * N + Cs * P + (Cs + Cg) * P)
    bb = 4.0 * S * S * ((Cs + Cg) * P + Cg * N + Cg * N + Cs * P)
    return sd, keep, fb, bb


def run_batch(dev, levels, F, dF, reps, flush):
    n = len(levels)
    descs = (L.OffkStencil * n)()
    ios = (L.OffkStencilIO * n)()
    fbytes = bbytes = 0.0
    for i, (sd, keep, fb, bb) in enumerate(levels):
        gd, dgd, w, bias, dw, db = keep
        Cg = sd.Cg
        C.memmove(C.byref(descs[i]), C.byref(sd), C.sizeof(sd))
        io = ios[i]
        io.g, io.d = gd.data_ptr(), gd.data_ptr() + 4 * Cg
        io.w, io.bias, io.out, io.dout = w.data_ptr(), bias.data_ptr(), F.data_ptr(), dF.data_ptr()
        io.dg, io.dd = dgd.data_ptr(), dgd.data_ptr() + 4 * Cg
        io.dg_fs = io.dd_fs = sd.g_fs
        io.dw, io.dbias = dw.data_ptr(), db.data_ptr()
        fbytes += fb
        bbytes += bb
    lib = L.lib()
    st = C.c_void_p(torch.cuda.current_stream().cuda_stream)
    out = {}
    for name, fn, nbytes in (("fwd", lib.offk_stencil_diff_fwd_batch, fbytes), ("bwd", lib.offk_stencil_diff_bwd_batch, bbytes)):
        for _ in range(2):
            L.check(fn(n, descs, ios, st), name)
        ts = []
        for _ in range(reps):
            flush()
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record()
            L.check(fn(n, descs, ios, st), name)
            b.record()
            torch.cuda.synchronize()
            ts.append(a.elapsed_time(b))
        ts.sort()
        ms = ts[len(ts) // 2]
        out[name] = {"MB": round(nbytes / 1e6, 1), "us": round(ms * 1e3, 1), "GBs": round(nbytes / ms / 1e6, 1)}
    return out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--sweep", action="store_true")
    ap.add_argument("--batch", type=int, default=48)
    ap.add_argument("--length", type=int, default=3)
    ap.add_argument("--reps", type=int, default=10)
    ap.add_argument("--drop", type=int, default=2, help="0 none, 2 seeded dropout (train mode)")
    ap.add_argument("--json", default=None)
    a = ap.parse_args()
    dev = torch.device("cuda")
    pk, src = peak()
    flush = Flusher(dev)
    res = {"peak_GBs": pk, "peak_source": src, "lib": L.LIB_PATH, "cases": []}
    B, Lg = a.batch, a.length
    P = B * (Lg - 1)
    stages = {"28": (28, 320, [0, 160]), "14": (14, 1056, [0, 160, 320, 480, 640]), "7": (7, 832, [0, 160])}
    tot = {"fwd": [0.0, 0.0], "bwd": [0.0, 0.0]}
    for st, (S, ctot, offs) in stages.items():
        F = torch.zeros(P, S, S, ctot, device=dev)
        dF = torch.randn(P, S, S, ctot, device=dev)
        levels = [make_level(dev, B, Lg, S, 128, 32, ctot, o, a.drop) for o in offs]
        r = run_batch(dev, levels, F, dF, a.reps, flush)
        for k in tot:
            tot[k][0] += r[k]["MB"]
            tot[k][1] += r[k]["us"]
        res["cases"].append({"case": f"stage{st} ({len(offs)} units, S={S}, B={B}, L={Lg})", **r})
        print(res["cases"][-1], flush=True)
        del F, dF, levels
    for k in tot:
        res[k + "_total"] = {"MB": round(tot[k][0], 1), "us": round(tot[k][1], 1), "GBs": round(tot[k][0] / tot[k][1] * 1e3, 1),
                             "frac_of_peak": round(tot[k][0] / tot[k][1] * 1e3 / pk, 3)}
        print(k, res[k + "_total"], flush=True)
    if a.sweep:
        for S in (28, 14, 7):
            for Cg in (128, 256, 512, 1024):
                Cs = 32
                per_clip = 4.0 * S * S * (Cg + Cs) * Lg * 2
                Bs = max(8, int(1.0e9 / per_clip))
                Ps = Bs * (Lg - 1)
                ctot = Cg + Cs
                F = torch.zeros(Ps, S, S, ctot, device=dev)
                dF = torch.randn(Ps, S, S, ctot, device=dev)
                lv = [make_level(dev, Bs, Lg, S, Cg, Cs, ctot, 0, a.drop)]
                r = run_batch(dev, lv, F, dF, max(3, a.reps // 2), flush)
                res["cases"].append({"case": f"sweep S={S} Cg={Cg} Cs={Cs} B={Bs} L={Lg}", **r,
                                     "fwd_frac": round(r["fwd"]["GBs"] / pk, 3), "bwd_frac": round(r["bwd"]["GBs"] / pk, 3)})
                print(res["cases"][-1], flush=True)
                del F, dF, lv
    if a.json:
        with open(a.json, "w") as f:
            json.dump(res, f, indent=1)


if __name__ == "__main__":
    main()
