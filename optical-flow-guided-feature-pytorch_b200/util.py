"""Mirror of the reference's util.py Sobel modules on the fused stencil kernel.

``SobelFilter(in, out)(x) -> (x_grad, y_grad)`` (util.py:20-50) and ``SobelFilter_Diagonal(in, out)(x)``
(util.py:52-77): depth-wise 3x3 cross-correlation, zero pad 1, frozen taps, no bias.  The reference modules take
NCHW tensors; the kernel works channels-last, so these stand-alone modules transpose at the boundary (inside the
OFF unit no transpose exists: the unit's GEMM epilogue already emits channels-last data).
"""
from __future__ import annotations

import ctypes as C

import torch
import torch.nn as nn

from . import _lib as L

SOBEL_X = [[1., 0., -1.], [2., 0., -2.], [1., 0., -1.]]       # util.py:29
SOBEL_Y = [[1., 2., 1.], [0., 0., 0.], [-1., -2., -1.]]       # util.py:30
SOBEL_DIAG = [[0., 1., 0.], [-1., 0., 1.], [0., -1., 0.]]     # util.py:61


def _stream(t):
    return C.c_void_p(torch.cuda.current_stream(t.device).cuda_stream)


def _desc(n, c_total, cs, k, h, w, out_ctot, out_coff):
    sd = L.OffkStencil()
    sd.B, sd.L, sd.Cg, sd.Cs, sd.K, sd.H, sd.W = n, 2, 0, cs, k, h, w     # L = 2, flat index: pair p reads frame p
    sd.g_fs = sd.d_fs = c_total * h * w
    sd.g_ps = sd.d_ps = c_total
    sd.out_ctot, sd.out_coff = out_ctot, out_coff
    sd.index_mode, sd.drop_mode, sd.keep_scale, sd.drop_p = L.INDEX_REFERENCE_FLAT, L.DROP_NONE, 1.0, 0.0
    return sd


class _DepthwiseStencil(torch.autograd.Function):
    """y[n, kk*C + c] = sum_ij w[c, kk, i, j] * x[n, c, y+i-1, x+j-1]   (x NCHW in, [N, K*C, H, W] out)."""

    @staticmethod
    def forward(ctx, x, w):
        if not x.is_cuda or x.dtype != torch.float32:
            raise RuntimeError("Sobel stencil: expects a CUDA fp32 tensor (there is no CPU fallback)")
        n, c, h, wd = x.shape
        k = w.shape[1]
        if c % 4:
            raise RuntimeError("Sobel stencil: channel count must be a multiple of 4")
        xl = x.permute(0, 2, 3, 1).contiguous()
        out = torch.empty(n, h, wd, k * c, device=x.device, dtype=x.dtype)
        lib, st = L.lib(), _stream(x)
        # the kernel caches <= 32 spatial channels (a power-of-two number of quads) per launch
        for kk in range(k):
            for c0 in range(0, c, 32):
                rem, done = min(32, c - c0), 0
                while done < rem:
                    cs_run = 1 << (min(32, rem - done).bit_length() - 1)
                    cc = c0 + done
                    sd = _desc(n, c, cs_run, 1, h, wd, k * c, kk * c + cc)
                    wk = w[cc:cc + cs_run, kk:kk + 1].contiguous()
                    L.check(lib.offk_stencil_diff_fwd(C.byref(sd), None, xl.data_ptr() + 4 * cc, wk.data_ptr(), None,
                                                      out.data_ptr(), st), "sobel_fwd")
                    done += cs_run
        ctx.save_for_backward(w)
        ctx.shape = (n, c, h, wd, k)
        return out.permute(0, 3, 1, 2)

    @staticmethod
    def backward(ctx, g):
        (w,) = ctx.saved_tensors
        n, c, h, wd, k = ctx.shape
        gl = g.permute(0, 2, 3, 1).contiguous()
        acc = torch.zeros(n, h, wd, c, device=g.device, dtype=g.dtype)
        tmp = torch.empty(2 * n, h, wd, c, device=g.device, dtype=g.dtype)
        lib, st = L.lib(), _stream(g)
        for kk in range(k):
            for c0 in range(0, c, 32):
                rem, done = min(32, c - c0), 0
                while done < rem:
                    cs_run = 1 << (min(32, rem - done).bit_length() - 1)
                    cc = c0 + done
                    sd = _desc(n, c, cs_run, 1, h, wd, k * c, kk * c + cc)
                    wk = w[cc:cc + cs_run, kk:kk + 1].contiguous()
                    L.check(lib.offk_stencil_diff_bwd(C.byref(sd), gl.data_ptr(), None, None, wk.data_ptr(), None, 0,
                                                      tmp.data_ptr() + 4 * cc, c * h * wd, None, None, st), "sobel_bwd")
                    done += cs_run
            acc += tmp[:n]
        return acc.permute(0, 3, 1, 2), None


class _FixedStencil(nn.Module):
    def __init__(self, input_dim, output_dim, kernels):
        super().__init__()
        if input_dim != output_dim:
            raise ValueError("depth-wise Sobel filters need input_dim == output_dim (groups=output_dim, util.py:36)")
        self.kernels = kernels

    def _taps(self, name, kernel, dim):
        conv = nn.Module()
        conv.weight = nn.Parameter(torch.tensor(kernel).expand(dim, 1, 3, 3).contiguous(), requires_grad=False)
        setattr(self, name, conv)


class SobelFilter(_FixedStencil):
    """util.py:20-50: returns (x_grad, y_grad); state_dict keys conv1.weight / conv2.weight."""

    def __init__(self, input_dim, output_dim):
        super().__init__(input_dim, output_dim, 2)
        self._taps("conv1", SOBEL_X, output_dim)
        self._taps("conv2", SOBEL_Y, output_dim)

    def forward(self, input):
        w = torch.cat([self.conv1.weight, self.conv2.weight], dim=1).to(input.device)    # [C, 2, 3, 3]
        y = _DepthwiseStencil.apply(input, w)
        c = input.shape[1]
        return y[:, :c], y[:, c:]


class SobelFilter_Diagonal(_FixedStencil):
    """util.py:52-77: state_dict key conv.weight."""

    def __init__(self, input_dim, output_dim):
        super().__init__(input_dim, output_dim, 1)
        self._taps("conv", SOBEL_DIAG, output_dim)

    def forward(self, input):
        return _DepthwiseStencil.apply(input, self.conv.weight.to(input.device))
