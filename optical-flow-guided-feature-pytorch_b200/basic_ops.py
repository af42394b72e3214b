"""Mirror of the reference's basic_ops.py (Identity, SegmentConsensus, ConsensusModule) on liboffk kernels.

The reference implements SegmentConsensus as a legacy (non-static) autograd.Function
(basic_ops.py:12-36), which raises on torch >= 1.5; the math is kept: forward
``mean(dim, keepdim=True)`` (:22), backward ``grad.expand(shape) / shape[dim]`` (:31).
"""
from __future__ import annotations

import ctypes as C

import torch

from . import _lib as L


class Identity(torch.nn.Module):
    """basic_ops.py:8-10"""

    def forward(self, input):
        return input


def _stream(t):
    return C.c_void_p(torch.cuda.current_stream(t.device).cuda_stream)


class SegmentConsensus(torch.autograd.Function):
    """avg consensus over dim=1 of a [B, T, C] CUDA tensor -> [B, 1, C]."""

    @staticmethod
    def forward(ctx, x, consensus_type="avg", dim=1):
        ctx.kind, ctx.shape = consensus_type, x.shape
        if consensus_type == "identity":
            return x
        if consensus_type != "avg":
            return None                                                     # basic_ops.py:25-26
        if dim != 1 or x.dim() != 3 or not x.is_cuda or x.dtype != torch.float32:
            raise RuntimeError("SegmentConsensus('avg'): expects a CUDA fp32 [B, T, C] tensor and dim=1 "
                               "(there is no CPU fallback)")
        x = x.contiguous()
        b, t, c = x.shape
        out = torch.empty(b, 1, c, device=x.device, dtype=x.dtype)
        L.check(L.lib().offk_segment_mean_fwd(x.data_ptr(), b, t, c, out.data_ptr(), _stream(x)), "segment_mean_fwd")
        return out

    @staticmethod
    def backward(ctx, g):
        if ctx.kind == "identity":
            return g, None, None
        b, t, c = ctx.shape
        g = g.contiguous()
        dx = torch.empty(b, t, c, device=g.device, dtype=g.dtype)
        L.check(L.lib().offk_segment_mean_bwd(g.data_ptr(), b, t, c, dx.data_ptr(), _stream(g)), "segment_mean_bwd")
        return dx, None, None


class ConsensusModule(torch.nn.Module):
    """basic_ops.py:38-46 ('rnn' maps to identity, :42)."""

    def __init__(self, consensus_type, dim=1):
        super().__init__()
        self.consensus_type = consensus_type if consensus_type != "rnn" else "identity"
        self.dim = dim

    def forward(self, input):
        if self.consensus_type == "identity":
            return input
        return SegmentConsensus.apply(input, self.consensus_type, self.dim)
