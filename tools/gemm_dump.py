"""Bring-up: run one single-k-block GEMM on a debug build (OFFK_DEBUG_DUMP) and decode CTA 0's stage-0 shared memory.
    OFFK_LIB=tools/_dbg/liboffk_dbg.so python tools/gemm_dump.py"""
import ctypes as C
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import off_b200  # noqa
from off_b200 import _lib as L, tables as T

sys.path.insert(0, os.path.join(ROOT, "tools"))
dev = torch.device("cuda")
lib = L.lib()
lib.offk_debug_set_dump.argtypes = [C.c_void_p]
dump = torch.zeros(64 * 1024, device=dev)
assert lib.offk_debug_set_dump(dump.data_ptr()) == 0


def swz_mn(c, k, atoms):
    kl, cc = k & 3, c & 7
    return (((k >> 2) * atoms + (c >> 3)) << 9) + (kl << 7) + (((((cc >> 1) ^ kl) << 1) | (cc & 1)) << 4)


def swz_k(r, c):
    return (r >> 3) * 1024 + (r & 7) * 128 + ((c ^ (r & 7)) << 4)


def go(M, N, K, a_mn, b_mn, spec_kind="dense"):
    torch.manual_seed(0)
    A = torch.randn(M, K, device=dev)
    B = torch.randn(N, K, device=dev)
    a_store = A.t().contiguous() if a_mn else A.contiguous()
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
    d.split_k = 1
    dump.zero_()
    L.check(lib.offk_gather_gemm(C.byref(d), L.PREC_TF32, None), "gemm")
    torch.cuda.synchronize()
    want = A.double() @ B.double().t()
    err = ((out.double() - want).norm() / want.norm()).item()
    sm = dump.cpu().numpy()
    bn = (N + 15) // 16 * 16
    atoms_b = (bn + 31) // 32
    Ah, Bh = A.cpu().numpy(), B.cpu().numpy()
    # decode A tile (rows 0..127 of CTA 0, k 0..31)
    badA = 0
    for m in range(min(M, 128)):
        for k in range(min(K, 32)):
            off = (swz_mn(m // 4, k, 4) + (m % 4) * 4) if a_mn else (swz_k(m, k // 4) + (k % 4) * 4)
            badA += sm[off // 4] != Ah[m, k]
    badB = 0
    for n in range(min(N, bn)):
        for k in range(min(K, 32)):
            off = (swz_mn(n // 4, k, atoms_b) + (n % 4) * 4) if b_mn else (swz_k(n, k // 4) + (k % 4) * 4)
            badB += sm[4096 + off // 4] != Bh[n, k]
    nzA = int((sm[:4096] != 0).sum())
    nzB = int((sm[4096:4096 + atoms_b * 1024] != 0).sum())
    print(f"M{M} N{N} K{K} A{'mn' if a_mn else 'k'} B{'mn' if b_mn else 'k'}: rel err {err:.3e} | smem A mismatches {badA} "
          f"(nonzero {nzA}/4096) | smem B mismatches {badB} (nonzero {nzB}/{atoms_b*1024}) | out nonzero {int((out != 0).sum())}", flush=True)
    return out, want


for M, N, K in [(128, 32, 32), (128, 64, 32)]:
    for a_mn, b_mn in [(False, False), (True, False), (False, True), (True, True)]:
        go(M, N, K, a_mn, b_mn)
