"""Clip-sharded data parallelism for the OFF path: one process per GPU, weights replicated, the only collective is
the gradient all-reduce of the OFF parameters (the reference has no distributed code; its nn.DataParallel use is
single-device).

The gradients already live in ONE flat fp32 buffer (engine.grads_flat), split into two buckets in the order the
backward pass finishes them: [stage convs + FC heads] first, [nine OFF units] last.  The first bucket is reduced on a
side stream while the unit gradients are still being computed.
"""
from __future__ import annotations

import torch
import torch.distributed as dist


class GradAllReducer:
    """Averages ``flat[lo:hi]`` ranges across ranks.  Device-agnostic (NCCL on CUDA, gloo on CPU tensors)."""

    def __init__(self, flat: torch.Tensor, ranges, group=None):
        self.flat, self.ranges, self.group = flat, list(ranges), group
        self.world = dist.get_world_size(group) if dist.is_initialized() else 1
        self._work = []

    def launch(self, i: int):
        if self.world == 1:
            return
        lo, hi = self.ranges[i]
        self._work.append((dist.all_reduce(self.flat[lo:hi], op=dist.ReduceOp.SUM, group=self.group, async_op=True), lo, hi))

    def finish(self):
        for work, lo, hi in self._work:
            work.wait()
            self.flat[lo:hi].mul_(1.0 / self.world)
        self._work.clear()


class DataParallelOFF:
    """Wraps an OFFEngine: ``step(g7, g14)`` = backward + overlapped gradient averaging.
    Each rank builds its engine with the LOCAL batch (clips are independent; because of the reference's flat-index
    quirk, parity is defined per rank against the reference run with batch = B/W on that rank's clips)."""

    def __init__(self, engine, group=None):
        self.engine = engine
        self.reducer = GradAllReducer(engine.grads_flat, [engine.stage_range, engine.unit_range], group)
        self.comm_stream = torch.cuda.Stream(device=engine.device) if engine.device.type == "cuda" else None

    def broadcast_parameters(self, src: int = 0):
        if self.reducer.world > 1:
            dist.broadcast(self.engine.params_flat, src=src, group=self.reducer.group)

    def backward(self, g7, g14):
        eng, red = self.engine, self.reducer

        def after_stage():
            if red.world == 1:
                return
            self.comm_stream.wait_stream(torch.cuda.current_stream(eng.device))
            with torch.cuda.stream(self.comm_stream):
                red.launch(0)

        eng.backward(g7, g14, after_stage=after_stage)
        if red.world > 1:
            red.launch(1)
            red.finish()
            torch.cuda.current_stream(eng.device).wait_stream(self.comm_stream)
        return eng.grads
