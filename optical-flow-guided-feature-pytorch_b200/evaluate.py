"""The reference's video-level evaluation protocol on the B200 path (test_rgb_off.py:83-236,
test_flow_off.py:313-396): 25 segments x 10 crops per video, the crops as the model's batch axis
(``batch=10, num_seg=25``, test_rgb_off.py:24-25,184), segment consensus inside the model, then per video

    rst = mean_crops(rst1) + 2 * mean_crops(rst2) + mean_crops(rst3)          (test_rgb_off.py:138)

with rst1 / rst2 / rst3 = the 7x7 OFF head, the backbone (RGB or Flow) score and the 14x14 OFF head; the prediction is the
arg-max, the metric the mean per-class accuracy (:210-219), and the raw per-video scores are written with
``np.savez(..., scores1=, scores2=, scores3=, label=)`` (:236) for the late fusion of score_fusion.ipynb.

Host-side protocol code only: the model call is ``BNInception_OFF.forward`` (Flow_OFF / RGB_OFF_v2 surface), which runs the
OFF section on liboffk.
"""
from __future__ import annotations

from typing import Iterable, Optional, Sequence

import numpy as np
import torch

NUM_CROPS = 10          # GroupOverSample(224, 256): 5 crops x 2 flips (test_rgb_off.py:45-47)
FUSE_WEIGHTS = (1.0, 2.0, 1.0)


def fuse_video_scores(rst1, rst2, rst3, weights: Sequence[float] = FUSE_WEIGHTS) -> np.ndarray:
    """[crops, classes] x 3 -> [1, classes]: crop mean per stream, weighted sum (test_rgb_off.py:138-141).  A stream may be
    None (no backbone score: OFF heads only)."""
    total = None
    for w, r in zip(weights, (rst1, rst2, rst3)):
        if r is None:
            continue
        r = np.asarray(r, dtype=np.float64)
        term = w * r.reshape(-1, r.shape[-1]).mean(axis=0)
        total = term if total is None else total + term
    return total.reshape(1, -1)


@torch.no_grad()
def eval_video(model, data, num_crops: int = NUM_CROPS):
    """One video.  ``data``: the crops' frames ``[num_crops * num_seg, C, 224, 224]`` ((crop, segment)-major, the
    ``data.view(25*10, 3, 224, 224)`` of test_rgb_off.py:184) for a model with a feature extractor, or the nine taps for a
    model without.  Returns (fused [1, classes], rst1, rst2, rst3) as numpy arrays (rst2 is None without a backbone)."""
    assert model.batch == num_crops, f"the model's batch axis carries the crops: build it with batch={num_crops}"
    was = model.training
    model.eval()
    try:
        if getattr(model, "modality_fuse", False):
            fused = model(data)                                   # one [crops, classes] tensor (test_rgb_off.py:156-165)
            r = fused.detach().float().cpu().numpy()
            return r.mean(axis=0).reshape(1, -1), r, None, None
        out = model(data)
        rst1, rst2, rst3 = out[0], out[1], out[2]
    finally:
        model.train(was)
    to_np = lambda t: None if t is None else t.detach().float().cpu().numpy().copy()
    rst1, rst2, rst3 = to_np(rst1), to_np(rst2), to_np(rst3)
    return fuse_video_scores(rst1, rst2, rst3), rst1, rst2, rst3


def per_class_accuracy(labels: Sequence[int], preds: Sequence[int], num_classes: Optional[int] = None):
    """Mean per-class accuracy from the confusion matrix (test_rgb_off.py:210-219), classes absent from ``labels`` skipped."""
    labels, preds = np.asarray(labels, dtype=np.int64), np.asarray(preds, dtype=np.int64)
    n = int(num_classes or max(labels.max(), preds.max()) + 1)
    cf = np.zeros((n, n), dtype=np.float64)
    np.add.at(cf, (labels, preds), 1.0)
    cnt, hit = cf.sum(axis=1), np.diag(cf)
    seen = cnt > 0
    acc = np.where(seen, hit / np.maximum(cnt, 1.0), np.nan)
    return float(np.nanmean(acc)), acc, cf


def evaluate(model, videos: Iterable, num_crops: int = NUM_CROPS, save_path: Optional[str] = None):
    """``videos`` yields (data, label).  Returns a dict with the mean per-class accuracy, the predictions and the raw score
    lists; ``save_path`` writes them in the reference's npz layout (test_rgb_off.py:236)."""
    s1, s2, s3, labels, preds = [], [], [], [], []
    for data, label in videos:
        fused, r1, r2, r3 = eval_video(model, data, num_crops)
        preds.append(int(np.argmax(np.mean(fused, axis=0))))      # test_rgb_off.py:205
        labels.append(int(label))
        s1.append(r1)
        s2.append(r2)
        s3.append(r3)
    acc, per_class, cf = per_class_accuracy(labels, preds)
    if save_path is not None:
        save_scores(save_path, s1, s2, s3, labels)
    return {"accuracy": acc, "per_class": per_class, "confusion": cf, "pred": preds, "label": labels,
            "scores1": s1, "scores2": s2, "scores3": s3}


def save_scores(path, scores1, scores2, scores3, labels):
    """np.savez('rgb_save_score_3', scores1=rst1_list, scores2=rst2_list, scores3=rst3_list, label=label_list)."""
    fill = lambda lst: [np.zeros_like(scores1[i]) if s is None else s for i, s in enumerate(lst)]
    np.savez(path, scores1=np.asarray(scores1), scores2=np.asarray(fill(scores2)), scores3=np.asarray(fill(scores3)),
             label=np.asarray(labels))
