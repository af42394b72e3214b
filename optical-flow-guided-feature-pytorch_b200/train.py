"""The training step around the OFF path (train_off.py:72,133-151) on the engine's flat buffers.

The reference does, per iteration: ``CrossEntropyLoss`` on each head against the clip labels repeated per frame pair
(train_off.py:133-146), ``clip_grad_norm(params, 20)`` (:149) and ``optim.Adam(lr=1e-3, betas=(0.9, 0.99),
weight_decay=5e-4).step()`` (:72,151) -- about a hundred small ATen launches over 108 parameter tensors.  Here the
parameters and gradients already live in two flat fp32 buffers, so the step is four liboffk launches: one fused
loss + dL/dlogits kernel per head, one sum-of-squares pass, one fused clip + Adam pass.  Nothing synchronises with the host:
the clip coefficient is computed on the device from the norm the previous kernel left in memory.
"""
from __future__ import annotations

import ctypes as C

import torch

from . import _lib as L
from . import spec as S


class FusedOFFTrainer:
    """``loss = trainer.loss_backward(fc7, fc14, target)`` then ``trainer.step()``.

    ``target`` holds one label per CLIP; rows of the per-pair logits ``[B*(L-1), 101]`` (RGB variant) use the label of their
    clip (``target.unsqueeze(1).repeat(1, L-1).view(-1)``, train_off.py:133); consensus outputs ``[B, 101]`` (Flow / v2)
    use it directly.  ``data_parallel`` (off_b200.dist.DataParallelOFF) averages the gradients across ranks inside
    ``loss_backward``; the update then runs on the averaged flat buffer.  Parameters without a gradient in the reference
    (``fc_action_motion_28``: never part of a returned output, RGB_OFF.py:787,860) are skipped by the update, exactly like
    ``optim.Adam`` skips ``grad is None``.
    """

    def __init__(self, net, lr=1e-3, betas=(0.9, 0.99), eps=1e-8, weight_decay=5e-4, max_norm=20.0, data_parallel=None):
        self.net, self.eng = net, net.engine
        self.lib = L.lib()
        self.lr, self.betas, self.eps, self.weight_decay, self.max_norm = lr, betas, eps, weight_decay, max_norm
        self.dp = data_parallel
        dev = self.eng.device
        self.exp_avg = torch.zeros_like(self.eng.params_flat)
        self.exp_avg_sq = torch.zeros_like(self.eng.params_flat)
        self.sumsq = torch.zeros(1, dtype=torch.float64, device=dev)
        self.loss = torch.zeros(1, dtype=torch.float32, device=dev)
        self.step_count = 0
        n_out = self.eng.B if self.eng.consensus else self.eng.P
        self.g7 = torch.zeros(n_out, S.NUM_CLASSES, device=dev)
        self.g14 = torch.zeros(n_out, S.NUM_CLASSES, device=dev)
        self.repeat = 1 if self.eng.consensus else self.eng.Lseg - 1
        # element ranges of the flat buffers that receive a gradient: everything but fc_action_motion_28.{weight, bias}
        lay = self.eng.layout
        lo28 = lay["fc_action_motion_28.weight"][0]
        hi28 = (lay["fc_action_motion_28.bias"][0] + S.NUM_CLASSES + 3) // 4 * 4
        ranges = [(0, lo28), (hi28, self.eng.n_flat)]
        self.ranges = [(a, b) for a, b in ranges if b > a]
        self._lo = (C.c_longlong * len(self.ranges))(*[a for a, _ in self.ranges])
        self._hi = (C.c_longlong * len(self.ranges))(*[b for _, b in self.ranges])

    def _stream(self):
        return C.c_void_p(torch.cuda.current_stream(self.eng.device).cuda_stream)

    def loss_backward(self, fc7, fc14, target):
        """Sum of the two heads' cross-entropy losses (mean over rows) + backward of the OFF section.  Returns the loss as a
        1-element device tensor (no synchronisation)."""
        assert target.dtype == torch.int64 and target.is_cuda and target.numel() * self.repeat == fc7.shape[0]
        st = self._stream()
        self.loss.zero_()
        for logits, grad in ((fc7, self.g7), (fc14, self.g14)):
            logits = logits.detach().contiguous()
            L.check(self.lib.offk_ce_loss_fwd_bwd(logits.data_ptr(), target.data_ptr(), logits.shape[0], logits.shape[1],
                                                  self.repeat, 1.0, self.loss.data_ptr(), grad.data_ptr(), st), "ce_loss")
        if self.dp is not None:
            self.dp.backward(self.g7, self.g14)
        else:
            self.eng.backward(self.g7, self.g14)
        return self.loss

    def step(self):
        """clip_grad_norm(max_norm) + Adam on the flat buffers (two launches)."""
        eng, st = self.eng, self._stream()
        self.step_count += 1
        self.sumsq.zero_()
        L.check(self.lib.offk_grad_sumsq(eng.grads_flat.data_ptr(), eng.n_flat, self.sumsq.data_ptr(), st), "grad_sumsq")
        L.check(self.lib.offk_clip_adam_step(eng.params_flat.data_ptr(), eng.grads_flat.data_ptr(), self.exp_avg.data_ptr(),
                                             self.exp_avg_sq.data_ptr(), self._lo, self._hi, len(self.ranges), self.lr,
                                             self.betas[0], self.betas[1], self.eps, self.weight_decay, self.step_count,
                                             self.max_norm, self.sumsq.data_ptr(), st), "clip_adam_step")

    def grad_norm(self):
        """Global gradient norm of the last step() (synchronises)."""
        return float(self.sumsq.sqrt().item())
