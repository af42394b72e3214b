"""The frozen TSN BN-Inception feature extractor that feeds the OFF path: it produces the nine Inception taps
ON THE DEVICE, under the reference's layer names, so that

  * ``BNInception_OFF(..., backbone="bninception")`` is a complete drop-in of the reference class: images in, the reference's
    return tuple out, ``state_dict()`` = the reference's 593 (RGB) / 576 (Flow, v2) entries, loadable from / savable to the
    reference's checkpoints key for key (RGB_OFF.py:43-263 decls, :362-594 forward; 10-channel stem Flow_OFF.py:53);
  * in deployment the 650 MB of taps per config-2 step never cross PCIe: 87 MB of images do (the end-to-end figure of bench.py
    with synthetic taps is the worst case of the synthetic-tap boundary, not of the product).

This is NOT part of the hot path this repository accelerates (the extractor is frozen, ``train_off.py:39-57``; every BASELINE
config feeds synthetic taps): it is a table-driven restatement on stock ``torch.nn.functional`` ops (cuDNN), run without
gradients.  Its arithmetic is pinned against the reference class by ``oracle/make_golden.py`` (fixture
``tests/golden/backbone_*.npz``) and its key / shape list by ``tests/golden/state_dict_keys_full_*.txt``.
"""
from __future__ import annotations

from collections import OrderedDict

import torch
import torch.nn as nn
import torch.nn.functional as F

# tag: (cin, 1x1, 3x3_reduce, 3x3, double_3x3_reduce, double_3x3, pool_proj, pool kind)      RGB_OFF.py:54-258
_NORMAL = OrderedDict([
    ("3a", (192, 64, 64, 64, 64, 96, 32, "avg")),
    ("3b", (256, 64, 64, 96, 64, 96, 64, "avg")),
    ("4a", (576, 224, 64, 96, 96, 128, 128, "avg")),
    ("4b", (576, 192, 96, 128, 96, 128, 128, "avg")),
    ("4c", (576, 160, 128, 160, 128, 160, 128, "avg")),
    ("4d", (608, 96, 128, 192, 160, 192, 128, "avg")),
    ("5a", (1056, 352, 192, 320, 160, 224, 128, "avg")),
    ("5b", (1024, 352, 192, 320, 192, 224, 128, "max")),
])
# stride-2 blocks: (cin, 3x3_reduce, 3x3, double_3x3_reduce, double_3x3)                      RGB_OFF.py:97-115,190-208
_REDUCE = {"3c": (320, 128, 160, 64, 96), "4e": (608, 128, 192, 192, 256)}
_ORDER = ["3a", "3b", "3c", "4a", "4b", "4c", "4d", "4e", "5a", "5b"]
TAPS = ["3a", "3b", "3c", "4a", "4b", "4c", "4d", "5a", "5b"]      # inception_X_output_out consumed by the OFF units


class BNInceptionBackbone(nn.Module):
    """forward(images [N, 3 | 10, 224, 224]) -> (taps {'3a'..'5b': [N, C, S, S]}, Feature_Generation_Score [N, num_classes],
    conv2_relu_3x3_out [N, 192, 56, 56] (the 4th output of RGB_OFF_v2.py:891))."""

    def __init__(self, num_classes=101, in_channels=3):
        super().__init__()
        self.num_classes = num_classes

        def cb(name, cin, cout, k, stride=1, pad=0):
            setattr(self, name, nn.Conv2d(cin, cout, k, stride, pad))
            setattr(self, name + "_bn", nn.BatchNorm2d(cout, eps=1e-5, momentum=0.9, affine=True))

        cb("conv1_7x7_s2", in_channels, 64, 7, 2, 3)
        cb("conv2_3x3_reduce", 64, 64, 1)
        cb("conv2_3x3", 64, 192, 3, 1, 1)
        for tag in _ORDER:
            p = f"inception_{tag}_"
            if tag in _NORMAL:
                cin, n1, r3, n3, rd, d, pp, _ = _NORMAL[tag]
                cb(p + "1x1", cin, n1, 1)
                cb(p + "3x3_reduce", cin, r3, 1)
                cb(p + "3x3", r3, n3, 3, 1, 1)
                cb(p + "double_3x3_reduce", cin, rd, 1)
                cb(p + "double_3x3_1", rd, d, 3, 1, 1)
                cb(p + "double_3x3_2", d, d, 3, 1, 1)
                cb(p + "pool_proj", cin, pp, 1)
            else:
                cin, r3, n3, rd, d = _REDUCE[tag]
                cb(p + "3x3_reduce", cin, r3, 1)
                cb(p + "3x3", r3, n3, 3, 2, 1)
                cb(p + "double_3x3_reduce", cin, rd, 1)
                cb(p + "double_3x3_1", rd, d, 3, 1, 1)
                cb(p + "double_3x3_2", d, d, 3, 2, 1)
        self.last_linear = nn.Linear(1024, num_classes)
        for prm in self.parameters():
            prm.requires_grad_(False)                 # train_off.py:39-46: only 'motion' parameters train
        self.eval()

    def train(self, mode=True):
        return super().train(False)                   # BatchNorm stays in eval mode (train_off.py:53-57)

    def _cbr(self, name, x):
        conv, bn = getattr(self, name), getattr(self, name + "_bn")
        return F.relu(bn(conv(x)), inplace=True)

    @torch.no_grad()
    def forward(self, x):
        x = self._cbr("conv1_7x7_s2", x)
        x = F.max_pool2d(x, 3, 2, 0, 1, ceil_mode=True)
        x = self._cbr("conv2_3x3_reduce", x)
        conv2 = self._cbr("conv2_3x3", x)
        x = F.max_pool2d(conv2, 3, 2, 0, 1, ceil_mode=True)
        taps = OrderedDict()
        for tag in _ORDER:
            p = f"inception_{tag}_"
            b3 = self._cbr(p + "3x3", self._cbr(p + "3x3_reduce", x))
            bd = self._cbr(p + "double_3x3_2", self._cbr(p + "double_3x3_1", self._cbr(p + "double_3x3_reduce", x)))
            if tag in _NORMAL:
                b1 = self._cbr(p + "1x1", x)
                if _NORMAL[tag][7] == "avg":
                    pool = F.avg_pool2d(x, 3, 1, 1, ceil_mode=True, count_include_pad=True)
                else:
                    pool = F.max_pool2d(x, 3, 1, 1, 1, ceil_mode=True)
                x = torch.cat([b1, b3, bd, self._cbr(p + "pool_proj", pool)], 1)
            else:
                x = torch.cat([b3, bd, F.max_pool2d(x, 3, 2, 0, 1, ceil_mode=True)], 1)
            if tag in TAPS:
                taps[tag] = x
        score = F.avg_pool2d(x, 7, 1, 0, ceil_mode=True, count_include_pad=True)
        score = self.last_linear(score.view(score.size(0), -1))
        return taps, score, conv2
