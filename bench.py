#!/usr/bin/env python
"""OFF-unit throughput benchmark (BASELINE.json metric: OFF-unit clips/sec, fwd+bwd; stencil HBM GB/s vs peak).

    python bench.py --gpus N --steps K --warmup W            # this repo's sm_100a path
    python bench.py --impl reference --gpus N ...            # the reference's CPU path (oracle port) on the host cores

A "step" is one forward+backward of the OFF sub-network (RGB_OFF.py:596-860) over one batch of synthetic
BN-Inception taps: config 2 of BASELINE.json (RGB variant, 48 clips x 3 segments per GPU, train-mode dropout,
cross-entropy on the 7x7 and 14x14 heads as in train_off.py:136-146).  For N > 1 every rank processes its own 48
clips (weak scaling) and the OFF-parameter gradients are averaged with one bucketed NCCL all-reduce per step.
Prints ONE JSON line on rank 0.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
for p in (ROOT, os.path.join(ROOT, "oracle")):
    if p not in sys.path:
        sys.path.insert(0, p)

METRIC = "off_unit_clips_per_sec_fwd_bwd"
# dram__bytes_read.sum + dram__bytes_write.sum of the three stencil_diff_fwd launches of one step, from the committed
# ncu --set full capture (profiles/); None until measured for the current kernel
NCU_TRAFFIC_BYTES = 286.8e6   # profiles/ncu_full_r01t_stencil_summary.txt: read 227.8 MB + write 59.0 MB (156 MB are written; the rest is still dirty in L2 at kernel end)
UNIT = "clips/s"


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--batch", type=int, default=48, help="clips per GPU")
    ap.add_argument("--length", type=int, default=3, help="segments per clip")
    ap.add_argument("--variant", default="rgb", choices=["rgb", "flow"])
    ap.add_argument("--precision", default="fp32", choices=["fp32", "tf32"],
                    help="fp32 = fp32-parity mode on the tensor cores (3xTF32, the BASELINE config-2 arithmetic); tf32 = one "
                         "kind::tf32 MMA per product (stated separately)")
    ap.add_argument("--cpu-clips", type=int, default=0, help="clips per CPU-baseline step (0 = --batch: the same batch as the GPU arm)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    return ap.parse_args()


def peaks():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            return json.load(f), "measured"
    except Exception:
        return {"hbm_gbs": 6650.0, "bf16_tflops": 1590.0}, "fallback"


# ----------------------------------------------------------------------------------------------- CPU arm
def cpu_oracle_rate(clips: int, length: int, variant: str, steps: int, warmup: int):
    """The reference's CPU path for this hot path (oracle port: the same ATen ops the reference modules call,
    RGB_OFF.py:596-860), fp32, all host threads, fwd+bwd with the same loss.  Returns (clips/s, s/step, threads)."""
    import torch
    import torch.nn.functional as F
    import off_oracle as O
    torch.set_num_threads(os.cpu_count() or 1)
    taps = O.make_taps(3, clips, length)
    prm = O.make_params(3, variant)
    masks = O.make_dropout_masks(3, clips, length)
    pairs = clips * (length - 1)
    tgt = torch.arange(pairs) % O.NUM_CLASSES
    if variant != "rgb":
        tgt = tgt[:clips]

    def loss(o):
        return F.cross_entropy(o["fc7"].reshape(len(tgt), -1), tgt) + F.cross_entropy(o["fc14"].reshape(len(tgt), -1), tgt)

    times = []
    for i in range(warmup + steps):
        t0 = time.perf_counter()
        O.off_forward_backward(taps, prm, clips, length, variant, masks, torch.float32, loss=loss)
        if i >= warmup:
            times.append(time.perf_counter() - t0)
    sec = sum(times) / len(times)
    return clips / sec, sec, torch.get_num_threads()


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    steps, warmup = max(1, min(args.steps, 6)), max(1, min(args.warmup, 1))
    clips = args.cpu_clips or args.batch
    rate, sec, threads = cpu_oracle_rate(clips, args.length, args.variant, steps, warmup)
    sample = (f"{clips} clips x {args.length} segments per step (one GPU's batch), fp32, {steps} timed steps after {warmup} "
              f"warm-up ({sec:.2f} s/step)")
    line = {
        "impl": "reference", "metric": METRIC, "value": rate, "unit": UNIT, "n_gpus": args.gpus, "steps": steps,
        "warmup": warmup, "ms_per_step": sec * 1e3, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "fp32", "data": "synthetic",
        "config": {"workload": f"{args.variant.upper()}_OFF OFF sub-network fwd+bwd, {args.batch} clips x {args.length} "
                               f"segments per GPU (BASELINE config 2); CPU arm: the same batch per step, a bounded number of steps",
                   "sample": sample},
        "cpu_baseline": {"value": rate, "unit": UNIT, "cores": threads, "kind": "port", "sample": sample},
        "e2e": {"value": rate, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line))


# ----------------------------------------------------------------------------------------------- clocks
class ClockSampler(threading.Thread):
    def __init__(self, index):
        super().__init__(daemon=True)
        self.index, self.samples, self.reasons, self.stop_flag = index, [], set(), False
        self.max_mhz = None

    def run(self):
        q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
             "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        while not self.stop_flag:
            try:
                out = subprocess.run(["nvidia-smi", f"--query-gpu={q}", "--format=csv,noheader,nounits", "-i",
                                      str(self.index)], capture_output=True, text=True, timeout=5).stdout.strip()
                f = [x.strip() for x in out.split(",")]
                self.samples.append(float(f[0]))
                self.max_mhz = float(f[1])
                for n, v in zip(names, f[2:]):
                    if v.lower().startswith("active"):
                        self.reasons.add(n)
            except Exception:
                pass
            time.sleep(0.1)

    def summary(self):
        s = sorted(self.samples)
        return {"sm_mhz": s[len(s) // 2] if s else None, "sm_max_mhz": self.max_mhz, "reasons": sorted(self.reasons)}


# ----------------------------------------------------------------------------------------------- GPU arm
def run_ours(args):
    import torch
    import torch.distributed as dist
    import torch.nn.functional as F
    import off_b200  # noqa: F401
    from off_b200 import spec as S
    from off_b200.modules import OFFSubNetwork
    from off_b200.dist import DataParallelOFF

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    B, Lg = args.batch, args.length
    net = OFFSubNetwork(B, Lg, args.variant, precision=args.precision, device=dev).train()
    eng = net.engine
    dp = DataParallelOFF(eng)
    dp.broadcast_parameters()
    torch.manual_seed(1234 + rank)
    taps_dev = net.tap_buffers()
    for t in taps_dev.values():
        t.copy_(torch.relu(torch.randn_like(t)))           # post-ReLU Inception taps (RGB_OFF.py:395)
    n_out = eng.P if not eng.consensus else B
    target = (torch.arange(n_out, device=dev) + rank) % S.NUM_CLASSES   # labels repeated per pair, train_off.py:133
    tap_bytes = sum(t.numel() * 4 for t in taps_dev.values())
    taps_host = {k: torch.empty(t.shape, dtype=t.dtype, pin_memory=True).copy_(t) for k, t in taps_dev.items()}
    loss_host = torch.empty((), dtype=torch.float32, pin_memory=True)

    def step(inputs=None):
        """One forward+backward.  inputs = None: taps already resident in HBM (device-timed `value`);
        inputs = a prefetch handle: the end-to-end path (taps staged from pinned host memory, loss read back)."""
        fc7, _, fc14 = net(taps_dev if inputs is None else inputs)
        loss = F.cross_entropy(fc7, target) + F.cross_entropy(fc14, target)
        # backward through the module's autograd.Function; the OFF-parameter gradients land in the flat buffer
        g7, g14 = torch.autograd.grad(loss, [fc7, fc14])
        dp.backward(g7, g14)
        if inputs is not None:
            loss_host.copy_(loss.detach(), non_blocking=True)
            torch.cuda.current_stream().synchronize()       # the user reads the loss every step
        return loss

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(n, e2e):
        barrier()
        t0, t1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        t0.record()
        if e2e:
            # public API: net.prefetch(host taps) starts the H2D copy of step i+1 into the idle input set while
            # step i computes; every step's 650 MB copy and 4-byte loss read-back are inside the timed region
            h = net.prefetch(taps_host)
            for i in range(n):
                h_next = net.prefetch(taps_host) if i + 1 < n else None
                step(h)
                h = h_next
        else:
            for _ in range(n):
                step()
        t1.record()
        barrier()
        ms = torch.tensor([t0.elapsed_time(t1)], device=dev)
        if world > 1:
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        return ms.item()

    for _ in range(max(3, args.warmup)):
        step()
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    ms = timed(args.steps, False)
    sampler.stop_flag = True
    timed(2, True)                                            # warm the staged path
    ms_e2e = timed(max(3, args.steps // 2), True) / max(3, args.steps // 2) * args.steps

    clips_total = B * world * args.steps
    value = clips_total / (ms * 1e-3)
    e2e_value = clips_total / (ms_e2e * 1e-3)

    # ---- roofline of the memory-bound kernel BASELINE.json names: the fused stencil (forward), one launch per
    # stage-fusion buffer (28: 3a,3b | 14: 3c..4d | 7: 5a,5b).  Timed live with CUDA events on the launching stream:
    # (1) cold, L2 flushed before every launch (the headline `achieved`), (2) in-step, events recorded around the
    # launches during extra forward passes (inputs just written by the unit GEMMs, partly L2-resident).
    pk, pk_src = peaks()
    roof = None
    if rank == 0:
        import ctypes as C
        flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
        flush_r = torch.zeros(256 << 20, dtype=torch.uint8, device=dev)
        flush_sink = torch.zeros((), dtype=torch.int64, device=dev)
        stream = C.c_void_p(torch.cuda.current_stream().cuda_stream)
        stage_levels = eng.stencil_fwd_tags                  # launch name -> the OFF units it serves
        lvl_bytes = {t: 4.0 * s * s * (S.GEN_C * eng.N + S.DOWN_C * eng.P + S.UNIT_C * eng.P)     # read G, read D, write M
                     for t, (cin, s) in S.LEVELS.items()}
        tot_ms, tot_bytes, per_stage = 0.0, 0.0, {}
        for st, step_fn in eng.stencil_fwd_steps.items():
            nbytes = sum(lvl_bytes[t] for t in stage_levels[st])
            reps, acc = 10, 0.0
            for _ in range(reps):
                # evict L2 (126 MB): write a 256 MB buffer, then read a second one so that the evicted lines are
                # written back BEFORE the timed launch (dirty lines would otherwise drain during it)
                flush.fill_(1)
                flush_sink.add_(flush_r.view(torch.int64).sum())
                a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                a.record()
                step_fn(stream)
                b.record()
                torch.cuda.synchronize()
                acc += a.elapsed_time(b)
            per_stage[st] = {"MB": round(nbytes / 1e6, 1), "us": round(acc / reps * 1e3, 1),
                             "GBs": round(nbytes / (acc / reps * 1e-3) / 1e9, 1)}
            tot_ms += acc / reps
            tot_bytes += nbytes
        achieved = tot_bytes / (tot_ms * 1e-3) / 1e9
        # backward: temporal half (streams G and dT, writes dG) and spatial half (dD, tap gradients) as two launches on
        # two streams, as the engine schedules them; algorithmic bytes 4*S^2*(160*P + 128*N + 128*N + 32*P) per level
        bwd_bytes = sum(4.0 * s * s * (S.UNIT_C * eng.P + 2 * S.GEN_C * eng.N + S.DOWN_C * eng.P) for _, s in S.LEVELS.values())
        side = torch.cuda.Stream(device=dev)
        n_lv = len(S.LEVELS)
        eng._set_dropout(True, None, 1)
        acc = 0.0
        for _ in range(10):
            flush.fill_(1)
            flush_sink.add_(flush_r.view(torch.int64).sum())
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            main = torch.cuda.current_stream()
            a.record(main)
            side.wait_event(a)
            eng.lib.offk_stencil_diff_bwd_batch_part(n_lv, eng._st_desc, eng._st_io, 1, C.c_void_p(main.cuda_stream))
            eng.lib.offk_stencil_diff_bwd_batch_part(n_lv, eng._st_desc, eng._st_io, 2, C.c_void_p(side.cuda_stream))
            main.wait_stream(side)
            b.record(main)
            torch.cuda.synchronize()
            acc += a.elapsed_time(b)
        bwd_us = acc / 10 * 1e3
        acc = 0.0
        for _ in range(10):                                  # same work as ONE launch (roles interleaved in one grid)
            flush.fill_(1)
            flush_sink.add_(flush_r.view(torch.int64).sum())
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record()
            eng.lib.offk_stencil_diff_bwd_batch(n_lv, eng._st_desc, eng._st_io, stream)
            b.record()
            torch.cuda.synchronize()
            acc += a.elapsed_time(b)
        fused_us = acc / 10 * 1e3
        bwd = {"MB": round(bwd_bytes / 1e6, 1), "us": round(fused_us, 1), "GBs": round(bwd_bytes / (fused_us * 1e-6) / 1e9, 1),
               "frac": round(bwd_bytes / (fused_us * 1e-6) / 1e9 / pk["hbm_gbs"], 3),
               "two_launch_us": round(bwd_us, 1),
               "note": "stencil_diff_bwd_kernel, all nine units in one launch (temporal and spatial blocks interleaved in one "
                       "grid), cold L2; two_launch_us = the halves as separate launches on two streams "
                       "(offk_stencil_diff_bwd_batch_part)"}
        # in-step: same launches timed inside full forward passes (single-stream issue so the events bracket them)
        eng.single_stream = True
        evs = []
        for _ in range(5):
            eng._set_dropout(True, None, 1)
            streams = eng._fork()
            for i, stp in enumerate(eng.fwd_sched.steps):
                hit = stp in eng.stencil_fwd_steps.values()
                if hit:
                    a = torch.cuda.Event(enable_timing=True); a.record()
                stp(C.c_void_p(streams[0].cuda_stream))
                if hit:
                    b = torch.cuda.Event(enable_timing=True); b.record(); evs.append((a, b))
        torch.cuda.synchronize()
        eng.single_stream = False
        in_step_ms = sum(a.elapsed_time(b) for a, b in evs) / 5
        roof = {"bound": "hbm", "kernel": f"stencil_diff_fwd_kernel ({len(per_stage)} launches per step: {' | '.join(per_stage)} stage units; 9 OFF units)",
                "achieved": achieved, "peak": pk["hbm_gbs"], "peak_source": pk_src, "unit": "GB/s",
                "frac": achieved / pk["hbm_gbs"], "traffic": NCU_TRAFFIC_BYTES, "bytes_per_step": tot_bytes,
                "per_stage": per_stage, "backward": bwd,
                "in_step": {"GBs": round(tot_bytes / (in_step_ms * 1e-3) / 1e9, 1), "us": round(in_step_ms * 1e3, 1),
                            "note": "same launches timed inside forward passes (inputs fresh from the unit GEMMs, partly L2-resident)"},
                "note": "algorithmic bytes 4*S^2*(128*N + 32*P + 160*P) per level (read G once, read D once, write the "
                        "160-channel slice once); `achieved` = cold launches, L2 flushed before each; traffic = ncu "
                        "dram__bytes_read+write summed over the forward stencil launches of one step (profiles/)"}

    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:      # the CPU baseline is timed at N = 1 only
        clips = args.cpu_clips or B
        rate, sec, threads = cpu_oracle_rate(clips, Lg, args.variant, 6, 1)
        cpu = {"value": rate, "unit": UNIT, "cores": threads, "kind": "port",
               "sample": f"{clips} clips x {Lg} segments (the GPU arm's batch), fwd+bwd, fp32, 6 timed steps after 1 warm-up "
                         f"({sec:.2f} s/step)"}

    if rank == 0:
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": max(3, args.warmup),
            "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": ("tf32 (tcgen05 kind::tf32 multiply, fp32 accumulate, fp32 storage)" if args.precision == "tf32" else
                      "fp32 (3xTF32: error-compensated tcgen05 kind::tf32 MMAs, fp32 accumulate, fp32 storage)"),
            "data": "synthetic",
            "config": {"workload": f"{args.variant.upper()}_OFF OFF sub-network fwd+bwd, {B} clips x {Lg} segments per GPU "
                                   f"(BASELINE config 2), train-mode dropout, CE loss on the 7x7 and 14x14 heads",
                       "clips_per_gpu": B, "segments": Lg, "precision": args.precision,
                       "l2": f"inputs larger than L2: {tap_bytes / 1e6:.0f} MB of taps per step vs 126 MB L2",
                       "parallelism": f"dp{world} (clip-sharded, NCCL all-reduce of {eng.n_flat} fp32 gradients)"},
            "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": tap_bytes, "d2h_bytes_per_step": 4,
                    "note": "taps copied from pinned host memory every step (copy of step i+1 overlaps compute of step i: "
                            "two input sets), loss read back every step"},
            "gpu_launches": (eng.launches_fwd + eng.launches_bwd) * args.steps,
            "clocks": sampler.summary(),
            "roofline": roof,
            "cpu_baseline": cpu,
            "flops_per_step": {"fwd": eng.flops_fwd, "bwd": eng.flops_bwd,
                               "tflops_achieved": (eng.flops_fwd + eng.flops_bwd) / (ms / args.steps * 1e-3) / 1e12},
        }
        print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    a = parse()
    if a.impl == "reference":
        run_reference(a)
    else:
        run_ours(a)
