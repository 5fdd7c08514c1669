"""CPU: the oracle restatements (oracle/*.py) against the committed fixtures generated FROM the reference's own functions
and modules (oracle/make_golden.py -> tests/golden/).  This is what pins the oracle."""
import json
import os

import numpy as np
import torch

from oracle import seg_model as ref_seg
from oracle import third_party as tp
from oracle import volume as ov

G = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def _idx():
    with open(os.path.join(G, "index_math.json")) as f:
        return json.load(f)


def test_index_math_matches_reference():
    idx = _idx()
    for c in idx["steps"]:
        assert ov.steps_for_sliding_window(c["image"], c["tile"], c["step"]) == c["out"], c
    for c in idx["n_slicers"]:
        sl = ov.sliding_window_slicers(c["image"], c["tile"])
        assert len(sl) == c["n"]
        assert [[[s.start, s.stop] for s in t[1:]] for t in sl[:3]] == c["first3"]
    for c in idx["find_integer_p"]:
        p = ov.find_integer_p(c["n"], c["s"])
        assert p == c["p"], c
        assert ov.calc_slices_to_crop(p, c["s"]) == c["crop"]
        assert ov.ideal_size(c["n"], c["s"]) == c["ideal"]
        assert ov.projected_size(c["n"], 0, c["s"]) == c["proj0"]
    for c in idx["get_pads"]:
        assert list(ov.get_pads(c["target"], c["d"])) == c["out"]
    for c in idx["get_patch"]:
        assert [[s.start, s.stop] for s in ov.get_patch_index(c["center"], c["size"])] == c["idx"]


def test_rotate_pad_fba_match_reference():
    z = np.load(os.path.join(G, "volume_ops.npz"))
    vol = torch.from_numpy(z["rot_in"])
    for a in (0, 90, -90, 180, -180, 270, -270, 360):
        assert np.array_equal(ov.rotate_vol_2d(vol, a).contiguous().numpy(), z[f"rot_{a}"]), a
    padded, pads = ov.target_pad(torch.from_numpy(z["pad_in"]), (10, 9, 7), mode="reflect")
    assert np.array_equal(padded.numpy(), z["pad_out"]) and np.array_equal(np.array(pads), z["pad_pads"])
    assert np.array_equal(ov.crop(padded, pads).numpy(), z["pad_crop"])
    vols = list(z["fba_in"])
    for key, p in (("fba_inf", "infinity"), ("fba_p2", 2), ("fba_p0", 0)):
        assert np.array_equal(ov.fba(vols, p), z[key]), key   # same numpy, same algorithm: bit-identical
    odd = list(z["fba_odd_in"])
    assert ov.fba(odd, "inf").shape == z["fba_odd_inf"].shape == (6, 5, 8)  # irfftn without `s` drops the odd last slice
    assert np.array_equal(ov.fba(odd, "inf"), z["fba_odd_inf"]) and np.array_equal(ov.fba(odd, "1"), z["fba_odd_p1"])


def test_gaussian_shim_matches_reference_call():
    z = np.load(os.path.join(G, "sliding_window.npz"))
    tp.compute_gaussian.cache_clear()
    g = tp.compute_gaussian((16, 16, 16), sigma_scale=1. / 8, value_scaling_factor=10, device=torch.device("cpu"))
    assert g.dtype == torch.float16 and np.array_equal(g.numpy(), z["gaussian_16"])
    assert float(g.max()) == 10.0 and float(g.min()) > 0.0


def _net(z, half):
    conv = torch.nn.Conv3d(1, 2, 3, padding=1)
    with torch.no_grad():
        conv.weight.copy_(torch.from_numpy(z["conv_w"]))
        conv.bias.copy_(torch.from_numpy(z["conv_b"]))
    conv.requires_grad_(False)
    return (lambda x: conv(x).half()) if half else conv


def test_sliding_window_blend_matches_reference():
    z = np.load(os.path.join(G, "sliding_window.npz"))
    data = torch.from_numpy(z["data"])
    for half in (0, 1):
        for gauss in (0, 1):
            sl = ov.sliding_window_slicers(data.shape[1:], [16, 16, 16])
            out = ov.sliding_window_logits(data.clone(), sl, _net(z, bool(half)), None, 1, [16, 16, 16], bool(gauss), False)
            want = z[f"logits_half{half}_gauss{gauss}"]
            assert out.dtype == torch.float16
            assert np.array_equal(out.numpy().view(np.uint16), want.view(np.uint16)), (half, gauss)  # bit-exact fp16


def test_oracle_segmodel_matches_reference_segmodel():
    z = np.load(os.path.join(G, "segmodel_tiny.npz"))
    m = ref_seg.build("tiny")  # same seed / construction order as the reference module in make_golden
    m.eval()
    assert sorted(m.state_dict().keys()) == list(z["keys"]) and len(m.state_dict()) == int(z["n_keys"])
    wsum = float(sum(p.detach().double().abs().sum() for p in m.parameters()))
    assert abs(wsum - float(z["weight_abs_sum"])) <= 1e-9 * wsum
    with torch.no_grad():
        out, up, skips = m(torch.from_numpy(z["x"]), return_inetermediate_feature=True)
    for got, key in ((out, "out"), (up, "up"), (skips[1], "skip1")):
        ref = torch.from_numpy(z[key])
        assert float((got - ref).norm() / ref.norm()) <= 1e-5, key   # fp32 bound of north_star


def test_oracle_flavr_matches_reference_flavr():
    """oracle/flavr.py against outputs of the reference's own UNet_3D_3D (both heads) and train_all.get_intermediate_features."""
    from oracle import flavr as of
    z = np.load(os.path.join(G, "flavr_small.npz"))
    x = torch.from_numpy(z["x"])
    for unc, tag in ((False, "plain"), (True, "uasr")):
        m = of.build(unc, seed=1234).eval()
        assert list(m.state_dict().keys()) == list(z[f"{tag}_keys"])
        wsum = float(sum(p.detach().double().abs().sum() for p in m.parameters()))
        assert abs(wsum - float(z[f"{tag}_weight_abs_sum"])) <= 1e-9 * wsum
        with torch.no_grad():
            xin = x.clone()
            out = m(xin)
        assert np.array_equal(xin.numpy(), z[f"{tag}_x_after"])          # in-place mean subtraction on the caller's tensor
        if unc:
            assert float((out[0] - torch.from_numpy(z["uasr_out"])).norm() / torch.from_numpy(z["uasr_out"]).norm()) <= 1e-5
            assert float((out[1] - torch.from_numpy(z["uasr_unc"])).norm() / torch.from_numpy(z["uasr_unc"]).norm()) <= 1e-5
        else:
            assert float((out - torch.from_numpy(z["plain_out"])).norm() / torch.from_numpy(z["plain_out"]).norm()) <= 1e-5
            with torch.no_grad():
                f = m(x.clone(), return_inetermediate_feature=True)
            assert np.allclose(f[1].numpy(), z["plain_x1"], rtol=1e-5, atol=1e-6) and f[4].shape == z["plain_x4"].shape
    from oracle.volume import zscore_normalization
    teacher = of.build(True, seed=1234).eval()
    img = torch.from_numpy(z["gif_img"]).clone()
    with torch.no_grad():
        feats = of.intermediate_features(teacher, img, torch.from_numpy(z["gif_lab"]), normalize=zscore_normalization)
    assert np.allclose(img.numpy(), z["gif_img_after"], rtol=1e-6, atol=1e-6)
    for i, key in ((1, "gif_f1"), (3, "gif_f3")):
        ref = torch.from_numpy(z[key])
        assert feats[i].shape == ref.shape and float((feats[i] - ref).norm() / ref.norm()) <= 1e-5


def test_oracle_sr_sweep_matches_reference_apply_to_vol_flavr():
    """tests/golden/sr_sweep.npz: the reference's own apply_to_vol_flavr (utils/sr_utils.py:102-135) on its own UNet_3D_3D
    (oracle/make_golden.py:sr_sweep_fixture) -- pad to a multiple of 16, Z-1 zero-padded windows, crop, transposed in-plane axes."""
    from oracle import flavr as of
    z = np.load(os.path.join(G, "sr_sweep.npz"))
    net = of.build(False, seed=1234).eval()
    with torch.no_grad():
        got = ov.apply_to_vol_flavr(net, torch.from_numpy(z["vol"]).clone())
    want = torch.from_numpy(z["out"])
    assert got.shape == want.shape == (16, 2, 24, 20)
    assert float((got - want).norm() / want.norm()) <= 1e-5


def test_oracle_wdsr_reproduces_the_reference_module():
    """oracle/wdsr.py vs tests/golden/wdsr_small.npz (the reference's own WDSR, `resize` by 1 = identity): same state_dict keys,
    same default init, same output and gradients."""
    import warnings
    from oracle import wdsr as ow
    z = np.load(os.path.join(G, "wdsr_small.npz"))
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        net = ow.build()
    assert list(net.state_dict().keys()) == list(z["keys"])
    assert np.array_equal(net.head.weight_v.detach().numpy(), z["head_v"])
    out = net(torch.from_numpy(z["x"]))
    assert float((out.detach() - torch.from_numpy(z["out"])).norm() / torch.from_numpy(z["out"]).norm()) < 1e-6
    (out * torch.from_numpy(z["cot"])).sum().backward()
    assert np.allclose(net.head.weight_g.grad.numpy(), z["grad_head_g"], rtol=1e-4, atol=1e-6)
    assert np.allclose(net.tail.conv0.weight_v.grad.numpy(), z["grad_tail_v"], rtol=1e-4, atol=1e-6)


def test_oracle_sr_sample_synthesis_matches_reference_getitem():
    """tests/golden/sr_degrade.npz: the reference's OWN `TrainSetMultiple.__getitem__` (utils/train_set.py:337-434) and `load_img`
    pre-filter (:321-333) with the resize stand-in injected, 40 seeded cases -> oracle.degrade.train_sample / blur_prefilter must
    reproduce them bit for bit (same statements, same `random` stream)."""
    import json
    import random
    from oracle import degrade as od
    z = np.load(os.path.join(G, "sr_degrade.npz"))
    cases = json.loads(bytes(z["cases"]).decode())
    assert len(cases) == 40
    image = np.concatenate([z["img"], z["lab"].astype(np.float32)], axis=-1)
    fx, fy = od.blur_prefilter(image, torch.from_numpy(z["kernel"]))
    assert np.array_equal(fx, z["filtered_x"]) and np.array_equal(fy, z["filtered_y"])
    branches = set()
    for c in cases:
        random.seed(c["seed"])
        lr, hr = od.train_sample(z["img"], z["lab"], fx, fy, c["patch_size"], c["slice_separation"], c["blur"], c["random_flip"])
        assert np.array_equal(lr.numpy(), z[c["key"] + "_lr"]), c
        assert np.array_equal(hr.numpy(), z[c["key"] + "_hr"]), c
        branches.add((lr.shape[1] < lr.shape[2], bool((lr[:, 0] == 0).all() or (lr[:, -1] == 0).all())))
    assert len(branches) >= 3      # both final permutations and the slice-dropout branch occur in the fixture


def test_resize_standin_properties():
    """The stand-in for the unavailable third-party `resize` (oracle/degrade.py): identity at step 1, exact on constants (the cubic
    weights sum to 1), nearest picks floor(p + 0.5), output length round(n / d), and it agrees with torch's own bicubic grid_sample
    (A = -0.75, align_corners=True on the same sample positions, border padding) away from the borders."""
    from oracle import degrade as od
    g = torch.Generator().manual_seed(0)
    x = torch.randn((2, 1, 40, 7), generator=g)
    assert torch.equal(od.resize_standin(x, (1, 1), 3), x)
    assert od.resize_standin(x, (4.0, 1), 3).shape == (2, 1, 10, 7) and od.resize_standin(x, (3.2, 1), 3).shape == (2, 1, 12, 7)
    c = torch.full((1, 1, 33, 3), 2.5)
    assert float((od.resize_standin(c, (2.5, 1), 3) - 2.5).abs().max()) < 1e-6
    r = torch.arange(40.0).reshape(1, 1, 40, 1)
    assert od.resize_standin(r, (4.0, 1), 0).flatten().tolist() == [2.0 + 4 * i for i in range(10)]     # p = 1.5, 5.5, ... -> floor(p + 0.5)
    # cubic vs torch.grid_sample on the same positions
    d = 4.0
    n_in, n_out = 40, 10
    pos = (torch.arange(n_out) + 0.5) * d - 0.5
    gy = 2 * pos / (n_in - 1) - 1
    gx = torch.linspace(-1, 1, 7)
    grid = torch.stack(torch.meshgrid(gy, gx, indexing="ij"), dim=-1).flip(-1).unsqueeze(0).repeat(2, 1, 1, 1)
    want = torch.nn.functional.grid_sample(x, grid, mode="bicubic", padding_mode="border", align_corners=True)
    got = od.resize_standin(x, (d, 1), 3)
    assert float((got[:, :, 1:-1] - want[:, :, 1:-1]).abs().max()) < 1e-5


def test_oracle_spatial_augmentation_matches_reference_augment_spatial():
    """tests/golden/spatial_aug.npz: the reference's OWN `augment_spatial` (utils/seg_utils.py:378-480; batchgenerators helpers
    restated, scipy's real map_coordinates), 15 seeded cases -> oracle.augment.augment_spatial_2d must reproduce them bit for bit."""
    import json
    from oracle import augment as oa
    z = np.load(os.path.join(G, "spatial_aug.npz"))
    cases = json.loads(bytes(z["cases"]).decode())
    assert len(cases) == 15
    for c in cases:
        np.random.seed(c["seed"])
        d, segs = oa.augment_spatial_2d(z["data"].copy(), [z["seg"].copy(), z["seg_sr"].copy(), z["uncertainty"].copy()],
                                        tuple(c["patch_size"]), p_scale_per_sample=c["p_scale"], p_rot_per_sample=c["p_rot"],
                                        enable_uncertainty=True)
        assert np.array_equal(d, z[c["key"] + "_data"]), c
        for s_, name in zip(segs, ("seg", "seg_sr", "uncertainty")):
            assert np.array_equal(s_, z[c["key"] + "_" + name]), (c, name)


def test_oracle_stage2_sample_matches_reference_getitem():
    """tests/golden/stage2_sample.npz: the reference's OWN `TrainSetMultipleSegSREfficient.__getitem__` (utils/train_set.py:102-159)
    with its own `MySpatialTransform` as `train_transform` (dummy-2D reshapes restated), 12 seeded cases -> oracle.augment.stage2_sample
    + spatial_only_transform must reproduce all four outputs bit for bit."""
    import json
    import random
    from oracle import augment as oa
    z = np.load(os.path.join(G, "stage2_sample.npz"))
    cases = json.loads(bytes(z["cases"]).decode())
    assert len(cases) == 12
    for c in cases:
        random.seed(c["seed"])
        np.random.seed(c["seed"])
        ps = c["patch_size"]
        tr = oa.spatial_only_transform((ps[2], ps[1], ps[0]), True, p_rot_per_sample=c["p_rot"], p_scale_per_sample=c["p_scale"])
        res = oa.stage2_sample(z["img"], z["lab"], z["unc"], ps, c["separation"], transform=tr)
        for name, v in zip(("img", "label_lr", "label", "uncertainty_lr"), res):
            assert np.array_equal(v.numpy(), z[c["key"] + "_" + name]), (c, name)
