"""GPU: SR-stage sample synthesis (rehrseg_b200/degrade.py) against tests/golden/sr_degrade.npz --
outputs of the reference's OWN `TrainSetMultiple.__getitem__` / `load_img` pre-filter (utils/train_set.py:321-434) with the
resize stand-in of oracle/degrade.py injected -- under the same `random` seeds, and the resampling kernel against the stand-in
definition on ragged shapes and non-integer steps."""
import json
import os
import random

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu
GOLD = os.path.join(os.path.dirname(__file__), "golden", "sr_degrade.npz")


def test_samples_match_the_reference_getitem():
    from rehrseg_b200 import degrade
    z = np.load(GOLD)
    cases = json.loads(bytes(z["cases"]).decode())
    assert len(cases) == 40
    kernel = torch.from_numpy(z["kernel"]).cuda()
    worst = 0.0
    for c in cases:
        ds = degrade.SRTrainSampler(c["patch_size"], c["slice_separation"], blur=c["blur"], random_flip=c["random_flip"],
                                    blur_kernel=kernel)
        ds.add_subject(z["img"], z["lab"])                       # blurred copies computed on the device (rehr_blur1d)
        random.seed(c["seed"])
        lr, hr = ds.sample(0)
        want_lr, want_hr = z[c["key"] + "_lr"], z[c["key"] + "_hr"]
        assert tuple(lr.shape) == want_lr.shape and tuple(hr.shape) == want_hr.shape, c
        assert np.array_equal(hr.cpu().numpy(), want_hr), c      # crop / pad / flips / permutation: exact
        err = float((lr.cpu() - torch.from_numpy(want_lr)).abs().max())
        worst = max(worst, err)
        assert err <= 1e-5, (c, err)                            # blur + cubic resampling in fp32
    print("worst |lr - reference|:", worst)


def test_prefilter_matches_load_img():
    from rehrseg_b200 import degrade
    z = np.load(GOLD)
    image = torch.cat((torch.from_numpy(z["img"]), torch.from_numpy(z["lab"]).float()), dim=-1).cuda()
    fx, fy = degrade.blur_prefilter(image, torch.from_numpy(z["kernel"]).cuda())
    assert float((fx.cpu() - torch.from_numpy(z["filtered_x"])).abs().max()) <= 1e-6
    assert float((fy.cpu() - torch.from_numpy(z["filtered_y"])).abs().max()) <= 1e-6


@pytest.mark.parametrize("shape,step,order", [((5, 1, 37, 19), 4.0, 3), ((3, 2, 64, 33), 3.2, 3), ((2, 1, 9, 8), 1.5, 3),
                                               ((4, 1, 40, 17), 4.0, 0), ((1, 1, 33, 5), 2.5, 0), ((2, 1, 16, 16), 0.5, 3)])
def test_resample_kernel_matches_the_standin(shape, step, order):
    from oracle import degrade as od
    from rehrseg_b200 import degrade
    g = torch.Generator().manual_seed(5)
    x = torch.randn(shape, generator=g)
    want = od.resize_standin(x, (step, 1), order=order)
    got = degrade.resize(x.cuda(), (step, 1), order=order)
    assert tuple(got.shape) == tuple(want.shape)
    tol = 0.0 if order == 0 else 2e-6
    assert float((got.cpu() - want).abs().max()) <= tol * max(1.0, float(want.abs().max()))
    # along Y too (the second step factor)
    want = od.resize_standin(x, (1, step), order=order)
    got = degrade.resize(x.cuda(), (1, step), order=order)
    assert float((got.cpu() - want).abs().max()) <= tol * max(1.0, float(want.abs().max()))


def test_batch_is_the_collated_samples():
    from rehrseg_b200 import degrade
    z = np.load(GOLD)
    ds = degrade.SRTrainSampler([32, 32, 1], 4.0, blur=True, random_flip=True, blur_kernel=torch.from_numpy(z["kernel"]).cuda())
    ds.add_subject(z["img"], z["lab"])
    random.seed(3)
    singles = [ds.sample(0) for _ in range(4)]
    random.seed(3)
    lr, hr = ds.batch([0, 0, 0, 0])
    assert lr.shape == (4, 2, 8, 32) and hr.shape == (4, 2, 32, 32)
    for k, (a, b) in enumerate(singles):
        assert torch.equal(lr[k], a) and torch.equal(hr[k], b)
