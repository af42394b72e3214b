"""Checkpoint compatibility with the reference's weight files.

The reference saves ``torch.save(net.state_dict(), '<timestamp>.pth')`` at the end of training (train_off.py:155-156,
train_off_consensus.py:232-233) and loads by merging a key-filtered subset of a checkpoint into the model's own
``state_dict`` (model_utils.py:192-216 ``'fc-action' not in k``, :238-261 ``'motion' not in k``, Flow_OFF.py:1398-1413
``k in model_state``).  Checkpoints written from ``nn.DataParallel`` carry a ``module.`` prefix (test_flow_off.py:52-58)
and TSN-wrapped ones a ``base_model.`` prefix (Flow_OFF.py:1399).  These helpers do the same on the B200 modules, whose
OFF ``state_dict`` keys equal the reference's.  Pure host code: no kernel is involved.
"""
from __future__ import annotations

import warnings
from collections import OrderedDict
from typing import Callable, Iterable, Optional

import torch


def normalize_keys(state_dict) -> "OrderedDict[str, torch.Tensor]":
    """Strip the wrappers' prefixes: 'module.' (DataParallel) and anything up to 'base_model.' (TSN wrapper)."""
    out = OrderedDict()
    for k, v in state_dict.items():
        k = k.split("base_model.")[-1]
        while k.startswith("module."):
            k = k[len("module."):]
        out[k] = v
    return out


def merge_state_dict(model: torch.nn.Module, checkpoint, keep: Optional[Callable[[str], bool]] = None,
                     strict_shapes: bool = True):
    """The reference's partial load: ``model_state.update({k: v for k, v in checkpoint.items() if keep(k)})`` followed by
    ``load_state_dict(model_state)``.  ``keep`` defaults to "every checkpoint key the model has" (Flow_OFF.py:1403);
    pass e.g. ``lambda k: 'motion' not in k`` (model_utils.py:240) to keep the model's fresh OFF branch.
    Returns (loaded keys, model keys the checkpoint did not provide, checkpoint keys that were ignored)."""
    ckpt = normalize_keys(checkpoint)
    state = model.state_dict()
    keep = keep or (lambda k: True)
    take = OrderedDict((k, v) for k, v in ckpt.items() if k in state and keep(k))
    wanted = [k for k in ckpt if keep(k)]
    if wanted and not take:
        raise RuntimeError(f"checkpoint merge loaded nothing: {len(wanted)} checkpoint keys pass the filter (e.g. "
                           f"{wanted[0]!r}) but none of them exists in the model (model keys look like {next(iter(state))!r})")
    absent = [k for k in wanted if k not in state]
    if absent:
        warnings.warn(f"{len(absent)} checkpoint keys pass the filter but are not in the model and were ignored "
                      f"(first: {absent[0]!r}); construct the model with its feature extractor (backbone=...) to load them")
    if strict_shapes:
        bad = [(k, tuple(v.shape), tuple(state[k].shape)) for k, v in take.items() if tuple(v.shape) != tuple(state[k].shape)]
        if bad:
            raise RuntimeError(f"checkpoint / model shape mismatch: {bad[:4]}")
    merged = OrderedDict(state)
    merged.update(take)
    model.load_state_dict(merged)
    missing = [k for k in state if k not in take]
    ignored = [k for k in ckpt if k not in take]
    return list(take), missing, ignored


def load_checkpoint(model: torch.nn.Module, path, keep: Optional[Callable[[str], bool]] = None, map_location="cpu"):
    """``torch.load`` + ``merge_state_dict``.  Reference files are plain state_dicts of tensors, so ``weights_only`` is on."""
    ckpt = torch.load(path, map_location=map_location, weights_only=True)
    if isinstance(ckpt, dict) and "state_dict" in ckpt and not torch.is_tensor(ckpt["state_dict"]):
        ckpt = ckpt["state_dict"]                      # util.save_checkpoint-style wrapper (util.py:6-10)
    return merge_state_dict(model, ckpt, keep)


def save_checkpoint(model: torch.nn.Module, path, data_parallel_prefix: bool = False,
                    only: Optional[Iterable[str]] = None):
    """train_off.py:155-156: ``torch.save(net.state_dict(), path)`` (tensors moved to the host).  ``data_parallel_prefix``
    writes the 'module.'-prefixed form a DataParallel-wrapped reference model would write."""
    sd = OrderedDict((k, v.detach().cpu().clone()) for k, v in model.state_dict().items() if only is None or k in set(only))
    if data_parallel_prefix:
        sd = OrderedDict(("module." + k, v) for k, v in sd.items())
    torch.save(sd, path)
    return list(sd)
