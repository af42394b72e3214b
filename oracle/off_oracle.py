"""CPU oracle for the OFF (Optical Flow guided Feature) hot path.

TEST INFRASTRUCTURE ONLY.  Nothing in the product package may import this
module; only ``tests/``, ``__graft_entry__.smoke()`` and the ``cpu_baseline`` /
``--impl reference`` legs of ``bench.py`` do, and there only as the checker or
as the timed CPU baseline -- never as the thing shipped.

It is a plain-PyTorch (CPU, fp32 or fp64) restatement of the data flow of
``/root/reference/RGB_OFF.py:596-860`` (learned depth-wise 3x3 spatial
gradient) and ``/root/reference/Flow_OFF.py:606-884`` /
``RGB_OFF_v2.py:613-891`` (fixed diagonal Sobel + segment consensus).  The
arithmetic itself lives in ATen (the reference has no kernels of its own,
SURVEY.md section 8c), so the oracle calls the same ATen operators
functionally, on explicit tensors, without the BN-Inception backbone.

Parity status: PINNED.  ``oracle/make_golden.py`` imports the real reference
classes from /root/reference, feeds them the same seeded taps / weights
through forward-pre-hooks and (a) asserts this
restatement matches them to fp32 round-off for forward and backward and
(b) writes the golden fixtures in ``tests/golden/`` that the CPU tests replay.
"""
from __future__ import annotations

import math
from collections import OrderedDict

import numpy as np
import torch
import torch.nn.functional as F

# ---------------------------------------------------------------------------
# Geometry of the nine OFF units (RGB_OFF.py:266-324 decls; taps at :395..:590)
# ---------------------------------------------------------------------------
LEVELS = OrderedDict([
    ("3a", (256, 28)), ("3b", (320, 28)), ("3c", (576, 14)),
    ("4a", (576, 14)), ("4b", (576, 14)), ("4c", (608, 14)), ("4d", (608, 14)),
    ("5a", (1024, 7)), ("5b", (1024, 7)),
])
GEN_C = 128      # motion_conv_gen_*   out channels (RGB_OFF.py:266)
DOWN_C = 32      # motion_spatial_down_* out channels (RGB_OFF.py:267)
NUM_CLASSES = 101  # fc_action_motion* (RGB_OFF.py:332-334)

SOBEL_X = [[1, 0, -1], [2, 0, -2], [1, 0, -1]]          # util.py:29
SOBEL_Y = [[1, 2, 1], [0, 0, 0], [-1, -2, -1]]          # util.py:30
SOBEL_DIAG = [[0, 1, 0], [-1, 0, 1], [0, -1, 0]]        # util.py:61


def param_shapes(variant: str = "rgb") -> "OrderedDict[str, tuple]":
    """Names/shapes of every OFF parameter, in declaration order.

    RGB_OFF.py:265-334 (``variant='rgb'``: includes motion_spatial_grad_*),
    Flow_OFF.py:275-345 (``variant='flow'``: no learned spatial grad).
    """
    s = OrderedDict()

    def conv(name, cout, cin, k, groups=1):
        s[name + ".weight"] = (cout, cin // groups, k, k)
        s[name + ".bias"] = (cout,)

    def unit(tag):
        cin = LEVELS[tag][0]
        conv("motion_conv_gen_" + tag, GEN_C, cin, 1)
        conv("motion_spatial_down_" + tag, DOWN_C, cin, 1)
        if variant == "rgb":
            conv("motion_spatial_grad_" + tag, DOWN_C, DOWN_C, 3, groups=DOWN_C)

    for t in ("3a", "3b", "3c"):
        unit(t)
    conv("motion_conv_trans_28", 64, 320, 7)
    conv("motion_conv1_trans_28a", 64, 64, 1)
    conv("motion_conv2_trans_28a", 64, 64, 3)
    conv("motion_conv3_trans_28a", 256, 64, 1)
    conv("motion_conv_branch_28a", 256, 64, 1)
    for b in ("28b", "28c"):
        conv("motion_conv1_trans_" + b, 64, 256, 1)
        conv("motion_conv2_trans_" + b, 64, 64, 3)
        conv("motion_conv3_trans_" + b, 256, 64, 1)
    for t in ("4a", "4b", "4c", "4d"):
        unit(t)
    conv("motion_conv_trans_14", 128, 1056, 5)
    conv("motion_conv1_trans_14a", 128, 128, 1)
    conv("motion_conv2_trans_14a", 128, 128, 3)
    conv("motion_conv3_trans_14a", 512, 128, 1)
    conv("motion_conv_expand_trans_14a", 512, 128, 1)
    conv("motion_conv1_trans_14b", 128, 512, 1)
    conv("motion_conv2_trans_14b", 128, 128, 3)
    conv("motion_conv3_trans_14b", 512, 128, 3)
    for t in ("5a", "5b"):
        unit(t)
    conv("motion_conv_trans", 256, 832, 3)
    conv("motion_conv1_trans", 256, 256, 1)
    conv("motion_conv2_trans", 256, 256, 3)
    conv("motion_conv3_trans", 1024, 256, 1)
    conv("motion_conv_branch_trans", 1024, 256, 1)
    s["fc_action_motion.weight"] = (NUM_CLASSES, 1024)
    s["fc_action_motion.bias"] = (NUM_CLASSES,)
    s["fc_action_motion_28.weight"] = (NUM_CLASSES, 256)
    s["fc_action_motion_28.bias"] = (NUM_CLASSES,)
    s["fc_action_motion_14.weight"] = (NUM_CLASSES, 512)
    s["fc_action_motion_14.bias"] = (NUM_CLASSES,)
    return s


# ---------------------------------------------------------------------------
# Platform-independent synthetic data (no torch RNG: a counter hash, so the
# same seed yields the same floats here, on the GPU box and in the fixtures).
# ---------------------------------------------------------------------------
def _hash_u01(seed: int, n: int, stream: int = 0) -> np.ndarray:
    x = np.arange(n, dtype=np.uint64)
    x += np.uint64((seed * 0x9E3779B97F4A7C15 + stream * 0xD1B54A32D192ED03) & 0xFFFFFFFFFFFFFFFF)
    with np.errstate(over="ignore"):
        x ^= x >> np.uint64(30)
        x *= np.uint64(0xBF58476D1CE4E5B9)
        x ^= x >> np.uint64(27)
        x *= np.uint64(0x94D049BB133111EB)
        x ^= x >> np.uint64(31)
    return ((x >> np.uint64(11)).astype(np.float64) + 0.5) / float(1 << 53)


def hash_uniform(seed: int, shape, lo=0.0, hi=1.0) -> torch.Tensor:
    n = int(np.prod(shape))
    u = _hash_u01(seed, n)
    return torch.from_numpy((lo + (hi - lo) * u).astype(np.float32)).reshape(shape)


def hash_normal(seed: int, shape) -> torch.Tensor:
    n = int(np.prod(shape))
    u1 = _hash_u01(seed, n, 1)
    u2 = _hash_u01(seed, n, 2)
    z = np.sqrt(-2.0 * np.log(u1)) * np.cos(2.0 * math.pi * u2)
    return torch.from_numpy(z.astype(np.float32)).reshape(shape)


def make_taps(seed: int, batch: int, length: int, dense: bool = False) -> "OrderedDict[str, torch.Tensor]":
    """Nine synthetic BN-Inception taps, fp32 NCHW, (b,t)-major frames.

    ``relu(randn)`` because the taps are post-ReLU concatenations
    (RGB_OFF.py:395); ``dense=True`` gives ``rand`` instead.
    """
    n = batch * length
    taps = OrderedDict()
    for i, (tag, (cin, s)) in enumerate(LEVELS.items()):
        shape = (n, cin, s, s)
        if dense:
            taps[tag] = hash_uniform(seed * 131 + i, shape)
        else:
            taps[tag] = torch.relu(hash_normal(seed * 131 + i, shape))
    return taps


def make_params(seed: int, variant: str = "rgb") -> "OrderedDict[str, torch.Tensor]":
    """Seeded weights distributed like the default nn.Conv2d / nn.Linear init
    (uniform(-1/sqrt(fan_in), 1/sqrt(fan_in)) for weight and bias)."""
    out = OrderedDict()
    fan = 1
    for i, (name, shape) in enumerate(param_shapes(variant).items()):
        if name.endswith(".weight"):
            fan = int(np.prod(shape[1:]))
        bound = 1.0 / math.sqrt(fan)
        out[name] = hash_uniform(seed * 7919 + 1000 + i, shape, -bound, bound)
    return out


def make_state_dict(seed: int, keys_shapes) -> "OrderedDict[str, torch.Tensor]":
    """Seeded tensors for an arbitrary ``(name, shape)`` list -- the FULL state_dict of a reference model (backbone
    included, tests/golden/state_dict_keys_full_*.txt): conv / linear weights and biases U(-1/sqrt(fan_in), 1/sqrt(fan_in)),
    BatchNorm weight and running_var U(0.5, 1.5), BatchNorm bias and running_mean U(-0.2, 0.2), num_batches_tracked 0."""
    out = OrderedDict()
    fan = 1
    for i, (name, shape) in enumerate(keys_shapes):
        s = seed * 104729 + 5000 + i
        leaf = name.rsplit(".", 1)[-1]
        if leaf == "num_batches_tracked":
            out[name] = torch.zeros((), dtype=torch.int64)
        elif "_bn." in name:
            out[name] = hash_uniform(s, shape, 0.5, 1.5) if leaf in ("weight", "running_var") else hash_uniform(s, shape, -0.2, 0.2)
        else:
            if leaf == "weight":
                fan = int(np.prod(shape[1:])) if len(shape) > 1 else fan
            bound = 1.0 / math.sqrt(max(fan, 1))
            out[name] = hash_uniform(s, shape, -bound, bound)
    return out


def make_dropout_masks(seed: int, batch: int, length: int, p_drop: float = 0.8):
    """Keep-masks (1 = keep) for the 12 dropout call sites of one forward
    (RGB_OFF.py:612..:827 spatial, :785,:791,:845 heads)."""
    pairs = batch * (length - 1)
    masks = OrderedDict()
    for i, (tag, (_, s)) in enumerate(LEVELS.items()):
        u = hash_uniform(seed * 977 + i, (pairs, DOWN_C, s, s))
        masks[tag] = (u >= p_drop).to(torch.uint8)
    for j, (tag, c) in enumerate((("fc28", 256), ("fc14", 512), ("fc7", 1024))):
        u = hash_uniform(seed * 977 + 100 + j, (pairs, c))
        masks[tag] = (u >= p_drop).to(torch.uint8)
    return masks


# ---------------------------------------------------------------------------
# tf32 operand emulation (for pinning the product's OFFK_PREC_TF32 mode)
# ---------------------------------------------------------------------------
# tcgen05.mma kind::tf32 reads the top 19 bits of every fp32 operand word (the low 13 mantissa bits are ignored) and
# accumulates the exact products in fp32.  ``mm="tf32_trunc"`` makes every dense contraction of the restatement (1x1 /
# KxK convs and their autograd: data gradient = dY x W, weight gradient = X x dY, bias gradient = sum of the
# truncated dY because the product folds it into the weight-gradient GEMM as an all-ones operand row) consume operands
# truncated the same way, everything else unchanged.  The depth-wise stencil, pools and dropout are not contractions.
def trunc_tf32(x: torch.Tensor) -> torch.Tensor:
    b = x.detach().to(torch.float32).contiguous().view(torch.int32) & -8192          # 0xFFFFE000
    return b.view(torch.float32).to(x.dtype)


class _Tf32Conv(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, w, b, stride, pad):
        ctx.save_for_backward(x, w)
        ctx.geom = (stride, pad, b is not None)
        return F.conv2d(trunc_tf32(x), trunc_tf32(w), b, stride, pad)

    @staticmethod
    def backward(ctx, dy):
        x, w = ctx.saved_tensors
        stride, pad, has_b = ctx.geom
        dyt = trunc_tf32(dy)
        dx = torch.nn.grad.conv2d_input(x.shape, trunc_tf32(w), dyt, stride, pad) if ctx.needs_input_grad[0] else None
        dw = torch.nn.grad.conv2d_weight(trunc_tf32(x), w.shape, dyt, stride, pad)
        db = dyt.sum((0, 2, 3)) if has_b else None
        return dx, dw, db, None, None


def _conv2d(x, w, b, stride=1, pad=0, mm="exact"):
    if mm == "exact":
        return F.conv2d(x, w, b, stride, pad)
    return _Tf32Conv.apply(x, w, b, stride, pad)


def _linear(x, w, b, mm="exact"):
    """The FC heads are not tensor-core contractions in the product (one fused pool + dropout + Linear kernel per head,
    plain fp32 FMAs, forward and backward): exact in every mode."""
    return F.linear(x, w, b)


# ---------------------------------------------------------------------------
# The restatement
# ---------------------------------------------------------------------------
def consensus_avg(x: torch.Tensor, dim: int = 1) -> torch.Tensor:
    """basic_ops.py:21-22 (forward) / :30-31 (backward = expand / shape[dim])."""
    return x.mean(dim=dim, keepdim=True)


def sobel_xy(x: torch.Tensor):
    """util.SobelFilter.forward (util.py:46-50): depth-wise, pad 1, no bias."""
    c = x.shape[1]
    wx = torch.tensor(SOBEL_X, dtype=x.dtype, device=x.device).expand(c, 1, 3, 3).contiguous()
    wy = torch.tensor(SOBEL_Y, dtype=x.dtype, device=x.device).expand(c, 1, 3, 3).contiguous()
    return F.conv2d(x, wx, None, 1, 1, 1, c), F.conv2d(x, wy, None, 1, 1, 1, c)


def sobel_diagonal(x: torch.Tensor) -> torch.Tensor:
    """util.SobelFilter_Diagonal.forward (util.py:74-77)."""
    c = x.shape[1]
    w = torch.tensor(SOBEL_DIAG, dtype=x.dtype, device=x.device).expand(c, 1, 3, 3).contiguous()
    return F.conv2d(x, w, None, 1, 1, 1, c)


def _drop(x, mask, p_drop):
    """nn.Dropout(p) in train mode with an injected keep mask: x*mask/(1-p)."""
    if mask is None:
        return x
    return x * mask.to(x.dtype).reshape(x.shape) * (1.0 / (1.0 - p_drop))


def _act(x, gates, site):
    """ReLU (RGB_OFF.py:340).  With ``gates`` (site -> bool tensor, e.g. the product's own activation signs) the ReLU is
    replaced by a multiplication with the given 0/1 pattern: identical wherever the pattern equals (x > 0), and at the
    rare element whose pre-activation sits within round-off of zero the forward changes by O(round-off) while the
    BACKWARD gate follows the supplied pattern.  This takes ReLU-gate flips -- a discontinuity, not an arithmetic
    error -- out of a gradient comparison."""
    if gates is None:
        return F.relu(x)
    return x * gates[site].to(x.dtype).reshape(x.shape)


def off_unit(tap, prm, tag, batch, length, variant, mask=None, p_drop=0.8, index_mode="reference_flat", mm="exact",
             gates=None):
    """One OFF unit, e.g. RGB_OFF.py:596-616 for level 3a.

    Returns (motion [P,160,S,S], gen_relu [N,128,S,S], down [P,32,S,S]).
    """
    s = tap.shape[-1]
    g = _act(_conv2d(tap, prm[f"motion_conv_gen_{tag}.weight"], prm[f"motion_conv_gen_{tag}.bias"], mm=mm),
             gates, "gen_" + tag)                                                                        # :597-598
    ch = g.shape[1]
    r = g.view(batch, -1, s, s)                       # :600
    temporal = (r[:, ch:] - r[:, :-ch]).reshape(-1, ch, s, s)  # :601-604
    pairs = batch * (length - 1)
    if index_mode == "reference_flat":
        spatial_frames = tap[:pairs]                  # :609  (flat-index quirk)
    else:                                             # 'aligned': frame (b,t) for pair (b,t)
        spatial_frames = tap.view(batch, length, *tap.shape[1:])[:, :-1].reshape(pairs, *tap.shape[1:])
    d = _conv2d(spatial_frames, prm[f"motion_spatial_down_{tag}.weight"], prm[f"motion_spatial_down_{tag}.bias"], mm=mm)  # :610
    if variant == "rgb":
        sg = F.conv2d(d, prm[f"motion_spatial_grad_{tag}.weight"], prm[f"motion_spatial_grad_{tag}.bias"], 1, 1, 1, DOWN_C)  # :611
    else:
        sg = sobel_diagonal(d)                        # Flow_OFF.py:622
    sg = _drop(sg, mask, p_drop)                      # :612
    return torch.cat((sg, temporal), dim=1), g, d     # :616


def off_forward(taps, prm, batch, length, variant="rgb", masks=None, p_drop=0.8,
                consensus=None, index_mode="reference_flat", mm="exact", gates=None):
    """OFF sub-network forward, RGB_OFF.py:596-860 / Flow_OFF.py:606-884.

    ``consensus``: None -> follow the variant (rgb: per-pair logits as in
    RGB_OFF.py:860; flow: segment average as in Flow_OFF.py:867-876).
    Returns a dict with the three heads and the stage-fusion intermediates.
    """
    if consensus is None:
        consensus = variant != "rgb"
    mk = (lambda k: masks[k]) if masks is not None else (lambda k: None)
    w = lambda n: prm[n + ".weight"]
    b = lambda n: prm[n + ".bias"]
    conv = lambda x, n, stride=1, pad=0: _conv2d(x, w(n), b(n), stride, pad, mm)
    relu = lambda x, site: _act(x, gates, site)          # ``site`` names the activation (see GATE_SITES)
    out = {}

    m = {t: off_unit(taps[t], prm, t, batch, length, variant, mk(t), p_drop, index_mode, mm, gates)[0] for t in LEVELS}

    # ---- resolution 28 (RGB_OFF.py:655-685)
    f28 = torch.cat((m["3a"], m["3b"]), 1)                               # :656
    t28 = conv(f28, "motion_conv_trans_28", 2, 3)                        # :657 (pre-ReLU kept for the branch, :665)
    r28 = relu(t28, "t28")                                               # :658
    h = relu(conv(r28, "motion_conv1_trans_28a"), "h1_28a")              # :659-660
    h = relu(conv(h, "motion_conv2_trans_28a", 1, 1), "h2_28a")          # :661-662
    h = conv(h, "motion_conv3_trans_28a")                                # :663
    s28 = relu(h + conv(t28, "motion_conv_branch_28a"), "s28a")          # :665-667
    for blk in ("28b", "28c"):                                           # :670-685
        h = relu(conv(s28, "motion_conv1_trans_" + blk), "h1_" + blk)
        h = relu(conv(h, "motion_conv2_trans_" + blk, 1, 1), "h2_" + blk)
        h = conv(h, "motion_conv3_trans_" + blk)
        s28 = relu(h + s28, "s" + blk)

    # ---- resolution 14 (RGB_OFF.py:759-780)
    f14 = torch.cat((m["3c"], m["4a"], m["4b"], m["4c"], m["4d"], s28), 1)  # :760
    t14 = relu(conv(f14, "motion_conv_trans_14", 2, 2), "t14")           # :762-763
    h = relu(conv(t14, "motion_conv1_trans_14a"), "h1_14a")
    h = relu(conv(h, "motion_conv2_trans_14a", 1, 1), "h2_14a")
    h = conv(h, "motion_conv3_trans_14a")
    s14 = relu(h + conv(t14, "motion_conv_expand_trans_14a"), "s14a")    # :769-771
    h = relu(conv(s14, "motion_conv1_trans_14b"), "h1_14b")
    h = relu(conv(h, "motion_conv2_trans_14b", 1, 1), "h2_14b")
    h = relu(conv(h, "motion_conv3_trans_14b", 1, 1), "h3_14b")          # 3x3 (:316) + ReLU before the add (:778)
    s14 = relu(s14 + h, "s14b")                                          # :779-780

    # ---- heads 28 / 14 (RGB_OFF.py:783-793)
    def head(x, fc, key):
        pooled = F.avg_pool2d(x, 7, stride=1, padding=0, ceil_mode=True, count_include_pad=True)  # global_pool :262
        pooled = _drop(pooled, mk(key), p_drop)
        pooled = torch.squeeze(pooled)                                   # :786 (drops the batch dim too when P == 1)
        return _linear(pooled, w(fc), b(fc), mm)

    p28 = F.max_pool2d(s28, 3, stride=2, dilation=1, ceil_mode=True)     # :353,:783
    fc28 = head(p28, "fc_action_motion_28", "fc28")
    fc14 = head(s14, "fc_action_motion_14", "fc14")

    # ---- resolution 7 (RGB_OFF.py:831-847)
    f7 = torch.cat((m["5a"], m["5b"], s14), 1)                           # :832
    t7 = relu(conv(f7, "motion_conv_trans", 1, 1), "t7")                 # :833-834
    h = relu(conv(t7, "motion_conv1_trans"), "h1_7")
    h = relu(conv(h, "motion_conv2_trans", 1, 1), "h2_7")
    h = conv(h, "motion_conv3_trans")
    s7 = h + conv(t7, "motion_conv_branch_trans")                        # :840-841 (no final ReLU)
    fc7 = head(s7, "fc_action_motion", "fc7")

    if consensus:                                                        # Flow_OFF.py:867-876
        seg = length - 1
        fc7 = consensus_avg(fc7.view(batch, seg, -1)).squeeze(1)
        fc28 = consensus_avg(fc28.view(batch, seg, -1)).squeeze(1)
        fc14 = consensus_avg(fc14.view(batch, seg, -1)).squeeze(1)
    out.update(fc7=fc7, fc28=fc28, fc14=fc14, fusion28=f28, fusion14=f14, fusion7=f7, sum7=s7)
    return out


# ReLU sites of the path, in forward order (the names off_forward passes to _act)
GATE_SITES = (["gen_" + t for t in LEVELS] +
              ["t28", "h1_28a", "h2_28a", "s28a", "h1_28b", "h2_28b", "s28b", "h1_28c", "h2_28c", "s28c",
               "t14", "h1_14a", "h2_14a", "s14a", "h1_14b", "h2_14b", "h3_14b", "s14b", "t7", "h1_7", "h2_7"])


def to_dtype(d, dtype):
    return OrderedDict((k, v.to(dtype)) for k, v in d.items())


def off_forward_backward(taps, prm, batch, length, variant="rgb", masks=None, dtype=torch.float32,
                         tap_grads=False, loss="sum", mm="exact", gates=None):
    """Forward + backward with ``loss = fc7.sum() + fc14.sum()``.

    Returns (outputs dict, grads dict name->tensor [, tap grads dict]).
    fc_action_motion_28.* gets no gradient (never returned, RGB_OFF.py:860).
    """
    prm = OrderedDict((k, v.detach().to(dtype).requires_grad_(True)) for k, v in prm.items())
    taps = OrderedDict((k, v.detach().to(dtype).requires_grad_(tap_grads)) for k, v in taps.items())
    out = off_forward(taps, prm, batch, length, variant, masks, mm=mm, gates=gates)
    if loss == "sum":
        l = out["fc7"].sum() + out["fc14"].sum()
    else:
        l = loss(out)
    l.backward()
    grads = OrderedDict((k, (v.grad if v.grad is not None else torch.zeros_like(v)).detach()) for k, v in prm.items())
    outs = {k: v.detach() for k, v in out.items()}
    if tap_grads:
        return outs, grads, OrderedDict((k, v.grad.detach()) for k, v in taps.items())
    return outs, grads
