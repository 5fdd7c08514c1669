"""GPU: stage-2 spatial augmentation (rehrseg_b200/augment.py) against tests/golden/spatial_aug.npz --
outputs of the reference's OWN `augment_spatial` (utils/seg_utils.py:378-480) with scipy's real map_coordinates under seeded
`np.random` -- and against scipy / the oracle directly on other shapes."""
import json
import os

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu
GOLD = os.path.join(os.path.dirname(__file__), "golden", "spatial_aug.npz")
TOL = 2e-4      # fp32 coefficients and weights against scipy's fp64 (relative to the largest magnitude of the result)


def _close(got, want, tol=TOL):
    return float((got.cpu().double() - torch.from_numpy(np.asarray(want)).double()).abs().max()) <= tol * max(1.0, float(np.abs(want).max()))


def test_matches_the_reference_augment_spatial():
    from rehrseg_b200 import augment
    z = np.load(GOLD)
    cases = json.loads(bytes(z["cases"]).decode())
    data, seg, seg_sr, unc = (torch.from_numpy(z[k]).cuda() for k in ("data", "seg", "seg_sr", "uncertainty"))
    moved = 0
    for c in cases:
        np.random.seed(c["seed"])
        d, segs = augment.augment_spatial(data, [seg, seg_sr, unc], tuple(c["patch_size"]), p_scale_per_sample=c["p_scale"],
                                          p_rot_per_sample=c["p_rot"], enable_uncertainty=True)
        want_d = z[c["key"] + "_data"]
        assert tuple(d.shape) == want_d.shape
        assert _close(d, want_d), c
        assert _close(segs[2], z[c["key"] + "_uncertainty"]), c
        for got, name in ((segs[0], "seg"), (segs[1], "seg_sr")):
            want = z[c["key"] + "_" + name]
            agree = float((got.cpu().numpy() == want).mean())
            assert agree >= 0.999, (c, name, agree)      # (an indicator interpolated to exactly 0.5 may round either way in fp32)
        moved += int(not np.array_equal(want_d, z["data"][:, :, :want_d.shape[2], :want_d.shape[3]]))
    assert moved >= 8       # the fixture really rotates / scales most cases


@pytest.mark.parametrize("shape", [(3, 2, 17, 40), (1, 1, 64, 2), (2, 5, 33, 31)])
def test_prefilter_matches_scipy_spline_filter(shape):
    import ctypes as C
    from scipy import ndimage as ndi
    from rehrseg_b200._lib import lib, ptr, check, stream_ptr
    g = torch.Generator().manual_seed(1)
    x = torch.randn(shape, generator=g)
    want = np.stack([[ndi.spline_filter(x[b, c].numpy().astype(np.float64), order=3, mode="mirror") for c in range(shape[1])] for b in range(shape[0])])
    xc = x.cuda().clone()
    b, c, X, Y = shape
    check(lib().rehr_bspline_prefilter_axis(ptr(xc), b * c, X, Y, stream_ptr()))
    check(lib().rehr_bspline_prefilter_axis(ptr(xc), b * c * X, Y, 1, stream_ptr()))
    torch.cuda.synchronize()
    assert _close(xc, want, 2e-6)


def test_dummy_2d_wrapper_matches_the_oracle():
    from oracle import augment as oa
    from rehrseg_b200 import augment
    rng = np.random.RandomState(9)
    b, z_lr, X, Y = 2, 4, 48, 40
    dd = {"data": rng.randn(b, 1, z_lr, X, Y).astype(np.float32),
          "seg": (rng.rand(b, 1, z_lr, X, Y) > 0.5).astype(np.float32),
          "seg_sr": (rng.rand(b, 1, 4 * z_lr, X, Y) > 0.5).astype(np.float32),
          "uncertainty": rng.rand(b, 1, z_lr, X, Y).astype(np.float32)}
    for seed in (0, 1, 2, 3, 4, 5):
        want = oa.spatial_transform_dummy_2d({k: v.copy() for k, v in dd.items()}, (z_lr, X, Y), rng=np.random.RandomState(seed))
        got = augment.spatial_transform_dummy_2d({k: torch.from_numpy(v).cuda() for k, v in dd.items()}, (z_lr, X, Y),
                                                  rng=np.random.RandomState(seed))
        for k in ("data", "uncertainty"):
            assert tuple(got[k].shape) == want[k].shape
            assert _close(got[k], want[k]), (seed, k)
        for k in ("seg", "seg_sr"):
            assert float((got[k].cpu().numpy() == want[k]).mean()) >= 0.999, (seed, k)


def test_stage2_samples_match_the_reference_getitem():
    """rehrseg_b200.augment.Stage2Sampler against tests/golden/stage2_sample.npz: outputs of the reference's OWN
    `TrainSetMultipleSegSREfficient.__getitem__` (utils/train_set.py:102-159) with its own `MySpatialTransform` as the transform,
    under the same `random` / `np.random` seeds."""
    import random
    from rehrseg_b200 import augment
    z = np.load(os.path.join(os.path.dirname(__file__), "golden", "stage2_sample.npz"))
    cases = json.loads(bytes(z["cases"]).decode())
    assert len(cases) == 12
    for c in cases:
        ds = augment.Stage2Sampler(c["patch_size"], c["separation"], p_rot_per_sample=c["p_rot"], p_scale_per_sample=c["p_scale"])
        ds.add_subject(z["img"], z["lab"], z["unc"])
        random.seed(c["seed"])
        np.random.seed(c["seed"])
        got = ds.sample(0)
        for name, g in zip(("img", "label_lr", "label", "uncertainty_lr"), got):
            want = z[c["key"] + "_" + name]
            assert tuple(g.shape) == want.shape, (c, name)
            if name in ("img", "uncertainty_lr"):
                assert _close(g, want, 3e-4), (c, name)
            else:
                assert float((g.cpu().numpy() == want).mean()) >= 0.999, (c, name)
