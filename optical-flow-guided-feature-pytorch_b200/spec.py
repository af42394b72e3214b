"""Static description of the OFF sub-network: levels, parameter names/shapes, flat-buffer layout.

Names and shapes are the reference's ``state_dict`` contract (RGB_OFF.py:265-334,
Flow_OFF.py:275-345).  All OFF parameters live in ONE flat fp32
buffer (and their gradients in a second one) so that the gradient all-reduce is a single
NCCL call and the unit's two 1x1 convs can run as one GEMM: for every level the
``motion_conv_gen_X`` and ``motion_spatial_down_X`` weights are adjacent (a [160, C_in]
matrix) and so are their biases.
"""
from __future__ import annotations

from collections import OrderedDict

LEVELS = OrderedDict([
    ("3a", (256, 28)), ("3b", (320, 28)), ("3c", (576, 14)),
    ("4a", (576, 14)), ("4b", (576, 14)), ("4c", (608, 14)), ("4d", (608, 14)),
    ("5a", (1024, 7)), ("5b", (1024, 7)),
])
GEN_C, DOWN_C, UNIT_C = 128, 32, 160
NUM_CLASSES = 101
DROP_P = 0.8  # nn.Dropout(p=0.8), RGB_OFF.py:356

# stage-fusion buffers: (total channels, spatial size, [(source, channel offset)])   RGB_OFF.py:656,760,832
STAGES = OrderedDict([
    ("28", (320, 28, [("3a", 0), ("3b", 160)])),
    ("14", (1056, 14, [("3c", 0), ("4a", 160), ("4b", 320), ("4c", 480), ("4d", 640), ("sum28c", 800)])),
    ("7", (832, 7, [("5a", 0), ("5b", 160), ("sum14b", 320)])),
])
LEVEL_STAGE = {"3a": "28", "3b": "28", "3c": "14", "4a": "14", "4b": "14", "4c": "14", "4d": "14", "5a": "7", "5b": "7"}

# (name, cout, cin, k, stride, pad) of every conv after the units, in forward order
STAGE_CONVS = [
    ("motion_conv_trans_28", 64, 320, 7, 2, 3),
    ("motion_conv1_trans_28a", 64, 64, 1, 1, 0), ("motion_conv2_trans_28a", 64, 64, 3, 1, 1),
    ("motion_conv3_trans_28a", 256, 64, 1, 1, 0), ("motion_conv_branch_28a", 256, 64, 1, 1, 0),
    ("motion_conv1_trans_28b", 64, 256, 1, 1, 0), ("motion_conv2_trans_28b", 64, 64, 3, 1, 1),
    ("motion_conv3_trans_28b", 256, 64, 1, 1, 0),
    ("motion_conv1_trans_28c", 64, 256, 1, 1, 0), ("motion_conv2_trans_28c", 64, 64, 3, 1, 1),
    ("motion_conv3_trans_28c", 256, 64, 1, 1, 0),
    ("motion_conv_trans_14", 128, 1056, 5, 2, 2),
    ("motion_conv1_trans_14a", 128, 128, 1, 1, 0), ("motion_conv2_trans_14a", 128, 128, 3, 1, 1),
    ("motion_conv3_trans_14a", 512, 128, 1, 1, 0), ("motion_conv_expand_trans_14a", 512, 128, 1, 1, 0),
    ("motion_conv1_trans_14b", 128, 512, 1, 1, 0), ("motion_conv2_trans_14b", 128, 128, 3, 1, 1),
    ("motion_conv3_trans_14b", 512, 128, 3, 1, 1),
    ("motion_conv_trans", 256, 832, 3, 1, 1),
    ("motion_conv1_trans", 256, 256, 1, 1, 0), ("motion_conv2_trans", 256, 256, 3, 1, 1),
    ("motion_conv3_trans", 1024, 256, 1, 1, 0), ("motion_conv_branch_trans", 1024, 256, 1, 1, 0),
]
CONV_BY_NAME = {c[0]: c for c in STAGE_CONVS}
FCS = [("fc_action_motion", 1024), ("fc_action_motion_28", 256), ("fc_action_motion_14", 512)]


def param_shapes(variant: str = "rgb") -> "OrderedDict[str, tuple]":
    """state_dict order of the reference declarations (what ``named_parameters`` would yield)."""
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
    for name, cout, cin, k, _, _ in STAGE_CONVS[:11]:
        conv(name, cout, cin, k)
    for t in ("4a", "4b", "4c", "4d"):
        unit(t)
    for name, cout, cin, k, _, _ in STAGE_CONVS[11:19]:
        conv(name, cout, cin, k)
    for t in ("5a", "5b"):
        unit(t)
    for name, cout, cin, k, _, _ in STAGE_CONVS[19:]:
        conv(name, cout, cin, k)
    for name, c in FCS:
        s[name + ".weight"] = (NUM_CLASSES, c)
        s[name + ".bias"] = (NUM_CLASSES,)
    return s


def flat_layout(variant: str = "rgb"):
    """name -> (offset, shape) inside the flat parameter buffer.  Offsets are multiples of 4 floats
    (16-byte aligned rows for the float4 weight loads).  Bucket order = reverse of the backward pass
    is obtained by iterating this dict backwards (heads/7-stage last)."""
    shapes = param_shapes(variant)
    order = []
    for tag in LEVELS:
        order += [f"motion_conv_gen_{tag}.weight", f"motion_spatial_down_{tag}.weight",
                  f"motion_conv_gen_{tag}.bias", f"motion_spatial_down_{tag}.bias"]
        if variant == "rgb":
            order += [f"motion_spatial_grad_{tag}.weight", f"motion_spatial_grad_{tag}.bias"]
    order += [n for n in shapes if n not in set(order)]
    assert sorted(order) == sorted(shapes)
    layout, off = OrderedDict(), 0
    pad_free = {f"motion_spatial_down_{t}.weight" for t in LEVELS} | {f"motion_spatial_down_{t}.bias" for t in LEVELS}
    for n in order:
        numel = 1
        for d in shapes[n]:
            numel *= d
        if n not in pad_free:
            off = (off + 3) // 4 * 4
        layout[n] = (off, shapes[n])
        off += numel
    return layout, (off + 3) // 4 * 4
