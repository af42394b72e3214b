"""Layout probe for the tcgen05 gather-GEMM: dense GEMMs with hand-made tables and structured operands, so that every
output element names the operand element the tensor core actually read.  python tools/gemm_probe.py
(bring-up tool for the MN-major shared-memory layouts; not a test)."""
import ctypes as C
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import off_b200  # noqa
from off_b200 import _lib as L, tables as T

dev = torch.device("cuda")


def run(M, N, K, A, B, a_mn, b_mn, split_k=1, tile_n=0):
    """A: [M,K] logical, B: [N,K] logical (torch fp32).  a_mn / b_mn: store the operand with m / n contiguous."""
    a_store = A.t().contiguous() if a_mn else A.contiguous()        # [K,M] or [M,K]
    b_store = B.t().contiguous() if b_mn else B.contiguous()
    idx = lambda off: np.array([(o, 0, 0) for o in off], dtype=T.IDX_DTYPE)
    spc = T.GemmSpec(
        M=M, N=N, K=K,
        a_row=idx(np.arange(M) * (1 if a_mn else K)), a_col=idx(np.arange(K) * (M if a_mn else 1)),
        b_row=(np.arange(N) * (1 if b_mn else K)).astype(np.int32), b_col=(np.arange(K) * (N if b_mn else 1)).astype(np.int32),
        out_row=(np.arange(M) * N).astype(np.int32), out_col=np.arange(N).astype(np.int32),
        a_mode=T.LOAD_VEC_ROW if a_mn else T.LOAD_VEC_K, b_mode=T.LOAD_VEC_ROW if b_mn else T.LOAD_VEC_K, out_vec=0)
    tabs = {k: torch.from_numpy(v).to(dev) for k, v in T.padded_tables(spc).items()}
    out = torch.zeros(M, N, device=dev)
    d = L.OffkGemm()
    d.M, d.N, d.K = M, N, K
    d.a_src, d.a_row, d.a_col = a_store.data_ptr(), tabs["a_row"].data_ptr(), tabs["a_col"].data_ptr()
    d.a_h, d.a_w, d.a_ones_row, d.a_mode = T.NO_BOX, T.NO_BOX, -1, spc.a_mode
    d.b_src, d.b_row, d.b_col, d.b_mode = b_store.data_ptr(), tabs["b_row"].data_ptr(), tabs["b_col"].data_ptr(), spc.b_mode
    d.out, d.out_row, d.out_col = out.data_ptr(), tabs["out_row"].data_ptr(), tabs["out_col"].data_ptr()
    d.split_k, d.tile_n, d.atomic_out = split_k, tile_n, int(split_k > 1)
    L.check(L.lib().offk_gather_gemm(C.byref(d), L.PREC_TF32, None), "gemm")
    torch.cuda.synchronize()
    return out


def report(name, got, want):
    bad = (got != want)
    nb = int(bad.sum())
    print(f"{name:60s} mismatches {nb}/{got.numel()}", flush=True)
    if nb:
        ii = bad.nonzero()[:6].tolist()
        print("     first:", [(m, n, float(got[m, n]), float(want[m, n])) for m, n in ii])
        rows = sorted(set(bad.nonzero()[:, 0].tolist()))
        cols = sorted(set(bad.nonzero()[:, 1].tolist()))
        print(f"     bad rows {rows[:12]}..{rows[-3:]} ({len(rows)})  bad cols {cols[:12]}..{cols[-3:]} ({len(cols)})")


def rnd(M, N, K, a_mn, b_mn):
    torch.manual_seed(0)
    A = torch.randn(M, K, device=dev)
    B = torch.randn(N, K, device=dev)
    got = run(M, N, K, A, B, a_mn, b_mn)
    want = A.double() @ B.double().t()
    return ((got.double() - want).norm() / want.norm()).item(), got.abs().max().item()


for (M, N, K) in [(128, 32, 32), (128, 160, 64), (980, 160, 64), (256, 160, 256), (128, 16, 32)]:
    for a_mn, b_mn in [(False, False), (True, False), (False, True), (True, True)]:
        e, mx = rnd(M, N, K, a_mn, b_mn)
        print(f"RANDOM M{M} N{N} K{K} A{'mn' if a_mn else 'k'} B{'mn' if b_mn else 'k'}: rel err {e:.3e}  max|got| {mx:.3f}", flush=True)

for (M, N, K) in [(128, 32, 32), (128, 160, 64)]:
    m = torch.arange(M, device=dev, dtype=torch.float32)
    n = torch.arange(N, device=dev, dtype=torch.float32)
    k = torch.arange(K, device=dev, dtype=torch.float32)
    for a_mn, b_mn in [(False, False), (True, False), (False, True), (True, True)]:
        tag = f"M{M} N{N} K{K} A{'mn' if a_mn else 'k'} B{'mn' if b_mn else 'k'}"
        # D[m,n] = A[m, n % K] when B[n,k] = delta(k == n % K)
        Bsel = (k[None, :] == (n[:, None] % K)).float()
        for what, A in (("A=m%251", (m[:, None] % 251).expand(M, K)), ("A=k%251", (k[None, :] % 251).expand(M, K))):
            got = run(M, N, K, A.contiguous(), Bsel, a_mn, b_mn)
            want = A[:, (torch.arange(N, device=dev) % K)]
            report(tag + " probeA " + what, got, want)
        # D[m,n] = B[n, m % K] when A[m,k] = delta(k == m % K)
        Asel = (k[None, :] == (m[:, None] % K)).float()
        for what, Bm in (("B=n", n[:, None].expand(N, K)), ("B=k%251", (k[None, :] % 251).expand(N, K))):
            got = run(M, N, K, Asel, Bm.contiguous(), a_mn, b_mn)
            want = Bm[:, (torch.arange(M, device=dev) % K)].t()
            report(tag + " probeB " + what, got, want)
