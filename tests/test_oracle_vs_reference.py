"""CPU, build container only: the oracle restatements against the LIVE reference functions imported unchanged from
/root/reference (skipped on the GPU box, where the tree does not exist)."""
import numpy as np
import pytest
import torch

from oracle import refimport
from oracle import volume as ov

pytestmark = pytest.mark.skipif(not refimport.available(), reason="/root/reference exists only in the build container")


def test_integer_helpers_exhaustive():
    po, pad, su = refimport.load("utils.patch_ops"), refimport.load("utils.pad"), refimport.load("utils.seg_utils")
    for n in range(1, 70):
        for s in (1.0, 1.25, 1.5, 2.0, 2.4, 3.0, 3.2, 4.0, 4.5, 6.0):
            assert ov.find_integer_p(n, s) == po.find_integer_p(n, s)
            assert ov.projected_size(n, 3, s) == po.projected_size(n, 3, s)
    for t in range(0, 40):
        for d in range(0, 40):
            assert ov.get_pads(t, d) == pad.get_pads(t, d)
    rng = np.random.default_rng(0)
    for _ in range(200):
        tile = [int(v) for v in rng.integers(4, 40, 3)]
        img = [t + int(v) for t, v in zip(tile, rng.integers(0, 90, 3))]
        step = float(rng.choice([0.25, 0.5, 0.75, 1.0]))
        assert ov.steps_for_sliding_window(img, tile, step) == su.compute_steps_for_sliding_window(img, tile, step)


def test_rotate_and_fba_random():
    rot, fb = refimport.load("utils.rotate"), refimport.load("utils.fba")
    g = torch.Generator().manual_seed(3)
    v = torch.randn((4, 6, 5), generator=g)
    for a in (0, 90, -90, 180, -180, 270, -270, 360):
        assert torch.equal(ov.rotate_vol_2d(v, a), rot.rotate_vol_2d(v, a))
    with pytest.raises(NotImplementedError):
        ov.rotate_vol_2d(v, 45)
    vols = [torch.randn((8, 6, 10), generator=g).numpy() for _ in range(4)]
    for p in ("infinity", "inf", 0.5, 2, "3"):
        assert np.array_equal(ov.fba(vols, p), fb.fba(vols, p))


def test_apply_to_vol_flavr_window_logic():
    """utils/sr_utils.py:102-135 hard-codes .cuda(); compare the restatement against a hand-built expectation on a
    model that tags every window (identity on the middle slices)."""
    calls = []

    def model(x):  # x [1, C, 4, Y, X]
        calls.append(x.clone())
        return x[:, :, 1:3].repeat_interleave(2, dim=2)  # [1, C, 4, Y, X]: four output slices per window

    img = torch.arange(5 * 1 * 20 * 18, dtype=torch.float32).reshape(5, 1, 20, 18)
    out = ov.apply_to_vol_flavr(model, img)
    assert out.shape == (16, 1, 18, 20) and len(calls) == 4  # in-plane axes come back transposed, as in the reference
    assert calls[0].shape == (1, 1, 4, 32, 32)
    assert float(calls[0][0, 0, 0].abs().sum()) == 0.0 and float(calls[-1][0, 0, 3].abs().sum()) == 0.0
    # window st uses slices st-1..st+2; output slice 4*st+k comes from input slice st (k<2) or st+1 (k>=2)
    for st in range(4):
        assert torch.equal(out[4 * st, 0], img[st, 0].T) and torch.equal(out[4 * st + 3, 0], img[st + 1, 0].T)
