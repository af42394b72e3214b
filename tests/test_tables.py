"""CPU: the gather-GEMM index tables reproduce conv2d forward / weight-grad / data-grad exactly."""
import numpy as np
import pytest
import torch
import torch.nn.functional as F

import off_b200  # noqa: F401
from off_b200 import tables as T

CASES = [
    # n, cin, h, w, cout, k, stride, pad, x_ctot, x_coff, y_ctot, y_coff
    (3, 8, 7, 7, 5, 1, 1, 0, 8, 0, 5, 0),        # unit 1x1
    (2, 6, 9, 8, 4, 3, 1, 1, 10, 3, 7, 2),       # 3x3 p1 on channel slices
    (2, 5, 14, 14, 6, 7, 2, 3, 5, 0, 6, 0),      # 7x7 s2 p3 (motion_conv_trans_28 geometry)
    (2, 4, 14, 14, 3, 5, 2, 2, 9, 5, 3, 0),      # 5x5 s2 p2 (motion_conv_trans_14 geometry)
    (4, 16, 1, 1, 7, 1, 1, 0, 16, 0, 7, 0),      # FC head as a 1x1 conv on 1x1 maps
]


CASES_NHWC = [
    (3, 8, 7, 7, 8, 1, 1, 0, 8, 0, 8, 0),
    (2, 8, 9, 8, 4, 3, 1, 1, 12, 4, 8, 4),
    (2, 4, 14, 14, 8, 7, 2, 3, 4, 0, 8, 0),
    (2, 4, 14, 14, 12, 5, 2, 2, 12, 8, 12, 0),
    (4, 16, 1, 1, 7, 1, 1, 0, 16, 0, 7, 0),
]


@pytest.mark.parametrize("case", CASES_NHWC)
@pytest.mark.parametrize("xl,yl", [("nhwc", "nhwc"), ("nchw", "nhwc")])
def test_tables_match_conv_nhwc(case, xl, yl):
    """Same, for the internal channels-last layout (and the NCHW-tap -> NHWC boundary of the unit conv);
    also checks the promises behind the vector load modes."""
    n, cin, h, w, cout, k, s, p, xct, xco, yct, yco = case
    g = T.ConvGeom(n, cin, h, w, cout, k, k, s, p, xct, xco, yct, yco)
    rng = np.random.default_rng(1)
    x_nchw = rng.standard_normal((n, xct, h, w))
    wt = rng.standard_normal((cout, cin, k, k))
    x = torch.tensor(x_nchw[:, xco:xco + cin], requires_grad=True)
    wtt = torch.tensor(wt, requires_grad=True)
    y = F.conv2d(x, wtt, None, s, p)
    dy_nchw = rng.standard_normal((n, yct, g.hout, g.wout))
    dy = torch.tensor(dy_nchw[:, yco:yco + cout])
    y.backward(dy)
    to = lambda a, lay: np.ascontiguousarray(a.transpose(0, 2, 3, 1)) if lay == "nhwc" else a
    back = lambda a, lay: a.transpose(0, 3, 1, 2) if lay == "nhwc" else a
    xbuf, dybuf = to(x_nchw, xl), to(dy_nchw, yl)
    w_l = np.ascontiguousarray(wt.transpose(0, 2, 3, 1)) if xl == "nhwc" else wt   # [cout,kh,kw,cin] for NHWC inputs

    spec = T.conv_fwd_spec(g, xl, yl)
    T.check_modes(spec)
    ybuf = np.zeros(to(np.zeros((n, yct, g.hout, g.wout)), yl).shape)
    T.scatter(spec, T.emulate(spec, xbuf, w_l), ybuf)
    np.testing.assert_allclose(back(ybuf, yl)[:, yco:yco + cout], y.detach().numpy(), atol=1e-10)

    spec = T.conv_wgrad_spec(g, xl, yl)
    T.check_modes(spec)
    dw = np.zeros_like(w_l)
    db = np.zeros(cout)
    T.scatter(spec, T.emulate(spec, xbuf, dybuf), dw, db, accumulate=True)
    dw_c = dw.transpose(0, 3, 1, 2) if xl == "nhwc" else dw
    np.testing.assert_allclose(dw_c, wtt.grad.numpy(), atol=1e-9)
    np.testing.assert_allclose(db, dy.sum((0, 2, 3)).numpy(), atol=1e-9)

    dxbuf = np.full(xbuf.shape, np.nan)
    for spec in T.conv_dgrad_specs(g, xl, yl):
        T.check_modes(spec)
        T.scatter(spec, T.emulate(spec, dybuf, w_l), dxbuf)
    got = back(dxbuf, xl)[:, xco:xco + cin]
    assert not np.isnan(got).any()
    np.testing.assert_allclose(got, x.grad.numpy(), atol=1e-9)


@pytest.mark.parametrize("case", CASES)
def test_tables_match_conv(case):
    n, cin, h, w, cout, k, s, p, xct, xco, yct, yco = case
    g = T.ConvGeom(n, cin, h, w, cout, k, k, s, p, xct, xco, yct, yco)
    rng = np.random.default_rng(0)
    xbuf = rng.standard_normal((n, xct, h, w))
    wt = rng.standard_normal((cout, cin, k, k))
    x = torch.tensor(xbuf[:, xco:xco + cin], requires_grad=True)
    wtt = torch.tensor(wt, requires_grad=True)
    y = F.conv2d(x, wtt, None, s, p)
    assert y.shape[2:] == (g.hout, g.wout)
    dybuf = rng.standard_normal((n, yct, g.hout, g.wout))
    dy = torch.tensor(dybuf[:, yco:yco + cout])
    y.backward(dy)

    # forward
    spec = T.conv_fwd_spec(g)
    D = T.emulate(spec, xbuf, wt)
    ybuf = np.zeros((n, yct, g.hout, g.wout))
    T.scatter(spec, D, ybuf)
    np.testing.assert_allclose(ybuf[:, yco:yco + cout], y.detach().numpy(), atol=1e-10)
    other = np.delete(ybuf, np.s_[yco:yco + cout], axis=1)
    assert not other.any()

    # weight + bias gradient
    spec = T.conv_wgrad_spec(g)
    D = T.emulate(spec, xbuf, dybuf)
    dw = np.zeros_like(wt)
    db = np.zeros(cout)
    T.scatter(spec, D, dw, db, accumulate=True)
    np.testing.assert_allclose(dw, wtt.grad.numpy(), atol=1e-9)
    np.testing.assert_allclose(db, dy.sum((0, 2, 3)).numpy(), atol=1e-9)

    # data gradient, one GEMM per stride-parity class; together they tile dX exactly once
    dxbuf = np.full((n, xct, h, w), np.nan)
    for spec in T.conv_dgrad_specs(g):
        D = T.emulate(spec, dybuf, wt)
        T.scatter(spec, D, dxbuf)
    got = dxbuf[:, xco:xco + cin]
    assert not np.isnan(got).any()
    np.testing.assert_allclose(got, x.grad.numpy(), atol=1e-9)
    assert np.isnan(np.delete(dxbuf, np.s_[xco:xco + cin], axis=1)).all()


@pytest.mark.parametrize("n,cin,h,cout,k,st,p", [(2, 8, 28, 4, 7, 2, 3), (3, 6, 14, 8, 5, 2, 2), (2, 4, 9, 3, 3, 1, 1), (1, 5, 11, 2, 4, 3, 1)])
def test_dgrad_class_as_stride1_correlation(n, cin, h, cout, k, st, p):
    """Host logic behind the TMA data gradients (OFFK_TGEMM_FREE_GEOM): every stride-parity class of dX equals a
    stride-1 correlation over dY with the class's taps in reverse order, the padding / output grid that
    conv_dgrad_specs records in ``extra`` and the weights gathered by dgrad_class_weight_index -- checked against
    autograd of conv2d."""
    g = T.ConvGeom(n, cin, h, h, cout, k, k, st, p)
    rng = np.random.default_rng(0)
    w = rng.standard_normal((cout, cin, k, k))
    dy = rng.standard_normal((n, cout, g.hout, g.wout))
    x = torch.zeros(n, cin, h, h, dtype=torch.float64, requires_grad=True)
    (torch.nn.functional.conv2d(x, torch.from_numpy(w), None, st, p) * torch.from_numpy(dy)).sum().backward()
    want = x.grad.numpy()
    got = np.zeros_like(want)
    seen = np.zeros((h, h), bool)
    for spc in T.conv_dgrad_specs(g, "nhwc", "nhwc", "nhwc"):
        ex = spc.extra
        assert ex["pad_h"] >= 0 and ex["pad_w"] >= 0
        D = T.emulate_dgrad_class(g, spc, dy.transpose(0, 2, 3, 1), w).reshape(n, ex["hc"], ex["wc"], cin)
        got[:, :, ex["a"]::st, ex["b"]::st] = D.transpose(0, 3, 1, 2)
        seen[ex["a"]::st, ex["b"]::st] = True
    assert seen.all()
    np.testing.assert_allclose(got, want, rtol=1e-10, atol=1e-10)
