"""CPU: the oracle restatement replays the golden vectors produced by the REAL reference (oracle/make_golden.py)."""
import glob
import os

import numpy as np
import pytest
import torch

import off_oracle as O

GOLD = os.path.join(os.path.dirname(__file__), "golden")
CASES = sorted(glob.glob(os.path.join(GOLD, "off_*.npz")))


def _digest_check(name, t, fix, key, rtol):
    t = t.detach().double().reshape(-1)
    scale = max(abs(float(fix[f"{key}.abssum"])) / t.numel(), 1e-30)
    assert abs(float(t.sum()) - float(fix[f"{key}.sum"])) <= rtol * max(float(fix[f"{key}.abssum"]), 1e-30), (name, key)
    assert abs(float(t.norm()) - float(fix[f"{key}.l2"])) <= rtol * max(float(fix[f"{key}.l2"]), 1e-30), (name, key)
    idx = torch.from_numpy(fix[f"{key}.idx"])
    np.testing.assert_allclose(t[idx].numpy(), fix[f"{key}.val"], rtol=0, atol=rtol * max(scale, np.abs(fix[f"{key}.val"]).max()))


@pytest.mark.parametrize("path", CASES, ids=[os.path.basename(p)[4:-4] for p in CASES])
def test_oracle_matches_reference_golden(path):
    fix = np.load(path)
    variant, batch, length, seed = str(fix["variant"]), int(fix["batch"]), int(fix["length"]), int(fix["seed"])
    taps = O.make_taps(seed, batch, length)
    prm = O.make_params(seed, variant)
    masks = O.make_dropout_masks(seed, batch, length) if int(fix["train"]) else None
    out, grads = O.off_forward_backward(taps, prm, batch, length, variant, masks, torch.float64)
    for k in ("fc7", "fc14", "fc28"):
        np.testing.assert_allclose(out[k].numpy().reshape(fix[k].shape), fix[k], rtol=0, atol=1e-9)
    for k in ("fusion28", "fusion14", "fusion7"):
        _digest_check(path, out[k], fix, k, 1e-9)
    for n, g in grads.items():
        _digest_check(path, g, fix, f"grad.{n}", 1e-8)


def test_sobel_known_answers():
    """util.py:29-30,61 known-answer vectors: ramp rows and a random tensor, from the reference classes."""
    fix = np.load(os.path.join(GOLD, "sobel_kat.npz"))
    x = torch.arange(5.0).repeat(5, 1).view(1, 1, 5, 5)
    gx, gy = O.sobel_xy(x)
    assert gx[0, 0, 2].tolist() == [-4.0, -8.0, -8.0, -8.0, 12.0]
    assert gy[0, 0, 0].tolist() == [-1.0, -4.0, -8.0, -12.0, -11.0] and gy[0, 0, 2].tolist() == [0.0] * 5
    assert O.sobel_diagonal(x)[0, 0, 2].tolist() == [1.0, 2.0, 2.0, 2.0, -3.0]
    np.testing.assert_array_equal(gx.numpy(), fix["ramp_gx"])
    xr = O.hash_normal(int(fix["rand_seed"]), (2, 4, 9, 7))
    rx, ry = O.sobel_xy(xr)
    np.testing.assert_allclose(rx.numpy(), fix["rand_gx"], atol=1e-6)
    np.testing.assert_allclose(ry.numpy(), fix["rand_gy"], atol=1e-6)
    np.testing.assert_allclose(O.sobel_diagonal(xr).numpy(), fix["rand_diag"], atol=1e-6)


def test_flat_index_quirk_is_reproduced():
    """RGB_OFF.py:609 slices the first B*(L-1) FLAT frames: for B > 1 spatial row p uses frame p, not frame (b,t)."""
    B, Lg = 2, 3
    taps = O.make_taps(7, B, Lg)
    prm = O.make_params(7, "rgb")
    a = O.off_unit(taps["5a"], prm, "5a", B, Lg, "rgb")[0]
    b = O.off_unit(taps["5a"], prm, "5a", B, Lg, "rgb", index_mode="aligned")[0]
    assert torch.equal(a[:2], b[:2])                       # pairs of clip 0 coincide
    assert not torch.allclose(a[2:, :32], b[2:, :32])      # spatial rows of clip 1 differ ...
    assert torch.equal(a[:, 32:], b[:, 32:])               # ... temporal rows never do
