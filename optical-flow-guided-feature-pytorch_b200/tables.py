"""Index tables that turn every dense op of the OFF path into one gather-GEMM.

``D[m,n] = sum_k A(m,k) * B(n,k)`` with
``A(m,k) = a_src[a_row[m].off + a_col[k].off]`` (zero outside the validity box),
``B(n,k) = b_src[b_row[n] + b_col[k]]`` and ``out[out_row[m] + out_col[n]]``
(see include/offk.h).  The tables depend only on the layer geometry, so they are
built once per (batch, length) plan on the host with numpy and uploaded; the
kernels contain no div/mod and no layout knowledge.

Forward  (nn.Conv2d call sites RGB_OFF.py:597..:847):  m = output pixel, k = (c,r,q), n = cout
Wgrad    (autograd of the same):  m = (c,r,q) [+ one "ones" row -> bias grad], k = output pixel, n = cout
Dgrad    : m = input pixel of one stride-parity class, k = (cout,r',q'), n = cin

``emulate()`` evaluates a spec with numpy so the tables can be verified on a CPU
against plain conv2d / autograd (tests/test_tables.py).
"""
from __future__ import annotations

from dataclasses import dataclass, field
from typing import Optional

import numpy as np

IDX_DTYPE = np.dtype([("off", "<i4"), ("y", "<i2"), ("x", "<i2")])


@dataclass
class ConvGeom:
    """y[n_img, cout, hout, wout] = conv2d(x[n_img, cin, hin, win], w[cout, cin, kh, kw], stride, pad);
    x / y may be channel slices [coff, coff+c) of wider NCHW buffers with ctot channels."""
    n_img: int
    cin: int
    hin: int
    win: int
    cout: int
    kh: int = 1
    kw: int = 1
    stride: int = 1
    pad: int = 0
    x_ctot: int = 0
    x_coff: int = 0
    y_ctot: int = 0
    y_coff: int = 0

    def __post_init__(self):
        self.x_ctot = self.x_ctot or self.cin
        self.y_ctot = self.y_ctot or self.cout
        self.hout = (self.hin + 2 * self.pad - self.kh) // self.stride + 1
        self.wout = (self.win + 2 * self.pad - self.kw) // self.stride + 1
        assert self.x_coff + self.cin <= self.x_ctot and self.y_coff + self.cout <= self.y_ctot
        lim = 2 ** 31 - 1
        assert self.n_img * self.x_ctot * self.hin * self.win < lim, "tensor too large for int32 offsets"
        assert self.n_img * self.y_ctot * self.hout * self.wout < lim, "tensor too large for int32 offsets"

    @property
    def kdim(self):
        return self.cin * self.kh * self.kw


@dataclass
class GemmSpec:
    """Host-side description of one gather-GEMM (numpy tables; pointers are bound later)."""
    M: int
    N: int
    K: int
    a_row: np.ndarray
    a_col: np.ndarray
    b_row: np.ndarray
    b_col: np.ndarray
    out_row: np.ndarray
    out_col: np.ndarray
    a_h: int = 0
    a_w: int = 0
    a_ones_row: int = -1
    a_klane: int = 0
    b_klane: int = 0
    b_dense: int = 0
    kind: str = ""
    extra: dict = field(default_factory=dict)


def _idx(off, y, x):
    t = np.empty(off.shape, IDX_DTYPE)
    t["off"] = off
    t["y"] = y
    t["x"] = x
    return np.ascontiguousarray(t.reshape(-1))


def _pixel_tables(g: ConvGeom):
    """Per output pixel m=(img,oh,ow): where its receptive field starts in x, and where it lives in y."""
    img, oh, ow = np.meshgrid(np.arange(g.n_img), np.arange(g.hout), np.arange(g.wout), indexing="ij")
    y0 = oh * g.stride - g.pad
    x0 = ow * g.stride - g.pad
    off = (img * g.x_ctot + g.x_coff) * (g.hin * g.win) + y0 * g.win + x0
    y_off = (img * g.y_ctot + g.y_coff) * (g.hout * g.wout) + oh * g.wout + ow
    return _idx(off, y0, x0), y_off.reshape(-1).astype(np.int32)


def _filter_tables(g: ConvGeom):
    """Per filter tap k=(c,r,q): offset inside the receptive field."""
    c, r, q = np.meshgrid(np.arange(g.cin), np.arange(g.kh), np.arange(g.kw), indexing="ij")
    return _idx(c * (g.hin * g.win) + r * g.win + q, r, q)


def conv_fwd_spec(g: ConvGeom) -> GemmSpec:
    """y = conv2d(x, w) + epilogue.  A = im2col(x) gathered on the fly, B = w viewed as [cout, cin*kh*kw]."""
    a_row, y_off = _pixel_tables(g)
    K = g.kdim
    box = not (g.kh == 1 and g.kw == 1 and g.pad == 0)
    return GemmSpec(
        M=g.n_img * g.hout * g.wout, N=g.cout, K=K,
        a_row=a_row, a_col=_filter_tables(g),
        b_row=(np.arange(g.cout) * K).astype(np.int32), b_col=np.arange(K, dtype=np.int32),
        out_row=y_off, out_col=(np.arange(g.cout) * (g.hout * g.wout)).astype(np.int32),
        a_h=g.hin if box else 0, a_w=g.win if box else 0,
        b_klane=1, b_dense=1 if K % 4 == 0 else 0, kind="fwd")


def conv_wgrad_spec(g: ConvGeom) -> GemmSpec:
    """dW[cout, (c,r,q)] += sum_pixels dY * im2col(x);  the extra all-ones A row yields db[cout].
    Rows are the filter taps (so the store into dW is contiguous along rows), k walks output pixels."""
    pix, y_off = _pixel_tables(g)
    K_w = g.kdim
    hw = g.hout * g.wout
    box = not (g.kh == 1 and g.kw == 1 and g.pad == 0)
    a_row = np.concatenate([_filter_tables(g), np.zeros(1, IDX_DTYPE)])
    img = np.repeat(np.arange(g.n_img), hw)
    b_col = (img * g.y_ctot * hw + np.tile(np.arange(hw), g.n_img)).astype(np.int32)
    return GemmSpec(
        M=K_w + 1, N=g.cout, K=g.n_img * hw,
        a_row=a_row, a_col=pix,
        b_row=((np.arange(g.cout) + g.y_coff) * hw).astype(np.int32), b_col=b_col,
        out_row=np.arange(K_w + 1, dtype=np.int32), out_col=(np.arange(g.cout) * K_w).astype(np.int32),
        a_h=g.hin if box else 0, a_w=g.win if box else 0,
        a_ones_row=K_w, a_klane=1, b_klane=1, kind="wgrad")


def conv_dgrad_specs(g: ConvGeom):
    """dX = conv_transpose(dY, w), one GEMM per stride-parity class (a,b) of the input pixel
    (ih = s*ih' + a):  only filter rows r = r0 + s*r' with r0 = (a+pad) % s reach that class, so the
    zero-stuffed taps of a strided transposed conv are never multiplied."""
    s, p = g.stride, g.pad
    hwo = g.hout * g.wout
    specs = []
    for a in range(s):
        for b in range(s):
            r0, q0 = (a + p) % s, (b + p) % s
            rs, qs = np.arange(r0, g.kh, s), np.arange(q0, g.kw, s)
            hc, wc = len(range(a, g.hin, s)), len(range(b, g.win, s))
            if hc == 0 or wc == 0:
                continue
            assert len(rs) > 0 and len(qs) > 0, "stride larger than kernel is not on the OFF path"
            dh, dw = (a + p - r0) // s, (b + p - q0) // s
            img, ih, iw = np.meshgrid(np.arange(g.n_img), np.arange(hc), np.arange(wc), indexing="ij")
            a_row = _idx((img * g.y_ctot + g.y_coff) * hwo + (ih + dh) * g.wout + (iw + dw), ih + dh, iw + dw)
            co, rr, qq = np.meshgrid(np.arange(g.cout), np.arange(len(rs)), np.arange(len(qs)), indexing="ij")
            a_col = _idx(co * hwo - rr * g.wout - qq, -rr, -qq)
            b_col = (co * (g.cin * g.kh * g.kw) + (r0 + s * rr) * g.kw + (q0 + s * qq)).reshape(-1).astype(np.int32)
            out_row = ((img * g.x_ctot + g.x_coff) * (g.hin * g.win) + (s * ih + a) * g.win + (s * iw + b))
            specs.append(GemmSpec(
                M=g.n_img * hc * wc, N=g.cin, K=g.cout * len(rs) * len(qs),
                a_row=a_row, a_col=a_col,
                b_row=(np.arange(g.cin) * (g.kh * g.kw)).astype(np.int32), b_col=b_col,
                out_row=out_row.reshape(-1).astype(np.int32),
                out_col=(np.arange(g.cin) * (g.hin * g.win)).astype(np.int32),
                a_h=g.hout, a_w=g.wout, kind=f"dgrad[{a},{b}]"))
    return specs


def emulate(spec: GemmSpec, a_src: np.ndarray, b_src: np.ndarray, a_relu: bool = False) -> np.ndarray:
    """Dense D[M,N] of a spec, evaluated with numpy (float64).  CPU-side check of the tables only."""
    a_src = a_src.reshape(-1).astype(np.float64)
    b_src = b_src.reshape(-1).astype(np.float64)
    ar, ac = spec.a_row, spec.a_col
    D = np.zeros((spec.M, spec.N))
    B = b_src[spec.b_row.astype(np.int64)[:, None] + spec.b_col.astype(np.int64)[None, :]]  # [N,K]
    step = max(1, (1 << 22) // max(spec.K, 1))
    for m0 in range(0, spec.M, step):
        r = ar[m0:m0 + step]
        off = r["off"].astype(np.int64)[:, None] + ac["off"].astype(np.int64)[None, :]
        if spec.a_h:
            y = r["y"].astype(np.int32)[:, None] + ac["y"].astype(np.int32)[None, :]
            x = r["x"].astype(np.int32)[:, None] + ac["x"].astype(np.int32)[None, :]
            ok = (y >= 0) & (y < spec.a_h) & (x >= 0) & (x < spec.a_w)
        else:
            ok = np.ones(off.shape, bool)
        A = np.where(ok, a_src[np.where(ok, off, 0)], 0.0)
        if a_relu:
            A = np.maximum(A, 0.0)
        if spec.a_ones_row >= 0 and m0 <= spec.a_ones_row < m0 + step:
            A[spec.a_ones_row - m0] = 1.0
        D[m0:m0 + step] = A @ B.T
    return D


def scatter(spec: GemmSpec, D: np.ndarray, out: np.ndarray, ones_out: Optional[np.ndarray] = None, accumulate=False):
    """Write / accumulate D through the output tables into the flat buffer ``out`` (numpy mirror of the epilogue
    without bias / activation)."""
    o = out.reshape(-1)
    rows = np.arange(spec.M)
    if spec.a_ones_row >= 0:
        if ones_out is not None:
            ones_out += D[spec.a_ones_row]
        rows = rows[rows != spec.a_ones_row]
    idx = spec.out_row.astype(np.int64)[rows][:, None] + spec.out_col.astype(np.int64)[None, :]
    if accumulate:
        np.add.at(o, idx, D[rows])
    else:
        o[idx] = D[rows]
    return out
