"""Clip-sharded data parallelism for the OFF path: one process per GPU, weights replicated, the only collective is
the gradient all-reduce of the OFF parameters (the reference has no distributed code; its nn.DataParallel use is
single-device).

The gradients already live in ONE flat fp32 buffer (engine.grads_flat), split into two buckets in the order the
backward pass finishes them: [stage convs + FC heads] first, [nine OFF units] last.  The first bucket is reduced on a
side stream while the unit gradients are still being computed.
"""
from __future__ import annotations

import ctypes
import os
import subprocess

import torch
import torch.distributed as dist


def _gpu_pci_bus_id(index: int):
    try:
        out = subprocess.run(["nvidia-smi", "--query-gpu=pci.bus_id", "--format=csv,noheader", "-i", str(index)],
                             capture_output=True, text=True, timeout=10).stdout.strip().splitlines()
        return out[0].strip().lower() if out else None
    except Exception:
        return None


def bind_to_gpu_numa_node(local_rank: int) -> dict:
    """Pin this process (CPU affinity + preferred memory node) to the NUMA node the GPU hangs off, BEFORE any pinned
    staging buffer is allocated: the host->device copies of the taps then read local DRAM and cross no socket link.
    One process per GPU, so the ranks spread over the nodes the way the GPUs do.  Best effort; returns what was done."""
    info = {"bound": False}
    bus = _gpu_pci_bus_id(local_rank)
    if not bus:
        info["reason"] = "no PCI bus id from nvidia-smi"
        return info
    bus = bus[-12:] if len(bus) > 12 else bus            # nvidia-smi prints an 8-digit domain, sysfs a 4-digit one
    try:
        node = int(open(f"/sys/bus/pci/devices/{bus}/numa_node").read().strip())
    except Exception as e:
        info["reason"] = f"numa_node of {bus}: {e!r}"[:120]
        return info
    info["node"] = node
    if node < 0:
        info["reason"] = "the platform reports no NUMA affinity for this GPU (single node or virtualised topology)"
        return info
    try:
        cpus = set()
        for part in open(f"/sys/devices/system/node/node{node}/cpulist").read().strip().split(","):
            lo, _, hi = part.partition("-")
            cpus |= set(range(int(lo), int(hi or lo) + 1))
        allowed = os.sched_getaffinity(0) & cpus
        if allowed:
            os.sched_setaffinity(0, allowed)
            info["cpus"] = len(allowed)
        # set_mempolicy(MPOL_PREFERRED, {node}): page allocations (incl. cudaHostAlloc'd staging) come from this node
        libc = ctypes.CDLL(None, use_errno=True)
        mask = (ctypes.c_ulong * 16)()
        mask[node // 64] = 1 << (node % 64)
        rc = libc.syscall(238, 1, ctypes.byref(mask), 16 * 64 + 1)       # __NR_set_mempolicy on x86-64, MPOL_PREFERRED = 1
        info["mempolicy"] = "preferred" if rc == 0 else f"errno {ctypes.get_errno()}"
        info["bound"] = bool(allowed)
    except Exception as e:
        info["reason"] = repr(e)[:120]
    return info


class GradAllReducer:
    """Averages ``flat[lo:hi]`` ranges across ranks.  Device-agnostic (NCCL on CUDA, gloo on CPU tensors).  On NCCL the
    1/W scaling happens inside the collective (ncclAvg): no extra pass over the gradients."""

    def __init__(self, flat: torch.Tensor, ranges, group=None):
        self.flat, self.ranges, self.group = flat, list(ranges), group
        self.world = dist.get_world_size(group) if dist.is_initialized() else 1
        self._avg = bool(dist.is_initialized() and flat.is_cuda and dist.get_backend(group) == "nccl")
        self._work = []

    def launch(self, i: int):
        if self.world == 1:
            return
        lo, hi = self.ranges[i]
        op = dist.ReduceOp.AVG if self._avg else dist.ReduceOp.SUM
        self._work.append((dist.all_reduce(self.flat[lo:hi], op=op, group=self.group, async_op=True), lo, hi))

    def finish(self):
        for work, lo, hi in self._work:
            work.wait()
            if not self._avg:
                self.flat[lo:hi].mul_(1.0 / self.world)
        self._work.clear()


class DataParallelOFF:
    """Wraps an OFFEngine: ``backward(g7, g14)`` = backward + overlapped gradient averaging.
    Each rank builds its engine with the LOCAL batch (clips are independent; because of the reference's flat-index
    quirk, parity is defined per rank against the reference run with batch = B/W on that rank's clips).

    Buckets follow the order in which the backward pass finishes the gradients: [stage convs + FC heads]
    as soon as the stage backward is done, then one bucket per OFF unit as its weight-gradient GEMM is issued (3a ... 5b in
    the plan's order; the unit's stencil gradients are complete by then).  Every bucket is reduced on a communication
    stream behind an event, so only the last unit's small all-reduce is exposed."""

    def __init__(self, engine, group=None, per_unit_buckets: bool = True, use_graphs: bool = False):
        self.engine = engine
        self.use_graphs = use_graphs          # single rank: replay the backward pass as one CUDA graph
        self.per_unit = per_unit_buckets
        self.tags = list(engine.unit_ranges)
        ranges = [engine.stage_range] + ([engine.unit_ranges[t] for t in self.tags] if per_unit_buckets else [engine.unit_range])
        self.reducer = GradAllReducer(engine.grads_flat, ranges, group)
        self.comm_stream = torch.cuda.Stream(device=engine.device) if engine.device.type == "cuda" else None

    def broadcast_parameters(self, src: int = 0):
        if self.reducer.world > 1:
            dist.broadcast(self.engine.params_flat, src=src, group=self.reducer.group)

    def _reduce_behind(self, stream, bucket):
        ev = torch.cuda.Event()
        ev.record(stream)
        self.comm_stream.wait_event(ev)
        with torch.cuda.stream(self.comm_stream):
            self.reducer.launch(bucket)

    def backward(self, g7, g14):
        eng, red = self.engine, self.reducer
        if red.world == 1:
            eng.backward(g7, g14, graph=self.use_graphs)
            return eng.grads

        def after_stage():
            self._reduce_behind(torch.cuda.current_stream(eng.device), 0)

        def after_unit(tag, stream):
            self._reduce_behind(stream, 1 + self.tags.index(tag))

        eng.backward(g7, g14, after_stage=after_stage, after_unit=after_unit if self.per_unit else None)
        if not self.per_unit:
            self._reduce_behind(torch.cuda.current_stream(eng.device), 1)
        with torch.cuda.stream(self.comm_stream):
            red.finish()
        torch.cuda.current_stream(eng.device).wait_stream(self.comm_stream)
        return eng.grads
