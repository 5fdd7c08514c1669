"""GPU: rotate / blur / FBA / mean fusion kernels (through the C-ABI) against the oracle and the reference fixtures.
Permutations are bit-exact; fp32 arithmetic is held to north_star's relative-L2 <= 1e-5."""
import os

import numpy as np
import pytest
import torch
import torch.nn.functional as F

pytestmark = pytest.mark.gpu
G = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def rel(a, b):
    a, b = np.asarray(a, dtype=np.float64), np.asarray(b, dtype=np.float64)
    return float(np.linalg.norm(a - b) / (np.linalg.norm(b) + 1e-30))


def test_rotate_vol_2d_bit_exact():
    from rehrseg_b200 import volume_ops as vo
    z = np.load(os.path.join(G, "volume_ops.npz"))
    vol = torch.from_numpy(z["rot_in"]).cuda()
    for a in (0, 90, -90, 180, -180, 270, -270, 360):
        got = vo.rotate_vol_2d(vol, a)
        assert np.array_equal(got.cpu().numpy(), z[f"rot_{a}"]), a           # reference fixture
    assert vo.rotate_vol_2d(vol, 0) is vol
    with pytest.raises(NotImplementedError):
        vo.rotate_vol_2d(vol, 30)
    g = torch.Generator().manual_seed(0)
    for shape, dt in (((33, 17, 5), torch.float32), ((64, 48), torch.float16), ((7, 9, 2, 3), torch.bfloat16), ((1, 5, 4), torch.float32),
                      ((512, 512, 12), torch.float32)):
        v = torch.randn(shape, generator=g).to(dt).cuda()
        for a in (90, -90, 180, 270):
            assert torch.equal(vo.rotate_vol_2d(v, a), torch.rot90(v, a // 90, [0, 1])), (shape, a)
        assert torch.equal(vo.rotate_vol_2d(vo.rotate_vol_2d(v, 90), -90), v)          # round trip
    e = torch.zeros((0, 4, 2)).cuda()
    assert vo.rotate_vol_2d(e, 90).shape == (4, 0, 2)                                   # empty input


def test_blur_matches_conv2d_same():
    from rehrseg_b200 import volume_ops as vo
    g = torch.Generator().manual_seed(1)
    for (Z, X, Y), L in (((5, 40, 33), 9), ((3, 16, 16), 1), ((2, 7, 50), 13), ((4, 32, 32), 4), ((1, 5, 3), 9)):
        x = torch.randn((Z, 1, X, Y), generator=g).cuda()
        k = torch.rand((1, 1, L, 1), generator=g).cuda()
        k = k / k.sum()
        want = F.conv2d(x, k, padding="same")
        got = vo.blur_along_x(x, k)
        assert got.shape == want.shape
        assert rel(got.cpu(), want.cpu()) <= 1e-5, (Z, X, Y, L)
    # linearity (size-independent property, full C5 plane size)
    x1, x2 = torch.randn((2, 1, 512, 512), generator=g).cuda(), torch.randn((2, 1, 512, 512), generator=g).cuda()
    k = torch.rand((1, 1, 9, 1), generator=g).cuda()
    lhs = vo.blur_along_x(2.0 * x1 + x2, k)
    rhs = 2.0 * vo.blur_along_x(x1, k) + vo.blur_along_x(x2, k)
    assert rel(lhs.cpu(), rhs.cpu()) <= 1e-5


def test_fba_matches_reference_fixture_and_oracle():
    from rehrseg_b200 import volume_ops as vo
    from oracle import volume as ov
    z = np.load(os.path.join(G, "volume_ops.npz"))
    vols = list(z["fba_in"])
    for key, p in (("fba_inf", "infinity"), ("fba_p2", 2), ("fba_p0", 0)):
        got = vo.fba(vols, p)
        assert isinstance(got, np.ndarray) and got.dtype == np.float32 and got.shape == z[key].shape
        assert rel(got, z[key]) <= 1e-5, (key, rel(got, z[key]))
    odd = list(z["fba_odd_in"])
    got = vo.fba(odd, "inf")
    assert got.shape == (6, 5, 8) and rel(got, z["fba_odd_inf"]) <= 1e-5          # odd last dim comes back one shorter
    assert rel(vo.fba(odd, "1"), z["fba_odd_p1"]) <= 1e-5
    # single volume: identity (up to FFT round-off); K identical volumes: identity for every p
    one = [vols[0]]
    assert rel(vo.fba(one, "infinity"), vols[0]) <= 1e-5
    assert rel(vo.fba([vols[0]] * 3, 2), vols[0]) <= 1e-5
    # larger random case against the numpy oracle, CUDA-tensor interface
    g = torch.Generator().manual_seed(2)
    big = [torch.randn((48, 40, 32), generator=g) for _ in range(4)]
    for p in ("infinity", 2.0):
        want = ov.fba([b.numpy() for b in big], p)
        got = vo.fba([b.cuda() for b in big], p)
        assert got.is_cuda and rel(got.cpu().numpy(), want) <= 1e-5, p
    with pytest.raises(ValueError):
        vo.fba([], "infinity")


def test_mean_fuse():
    from rehrseg_b200 import volume_ops as vo
    g = torch.Generator().manual_seed(3)
    vols = [torch.randn((17, 9, 5, 2), generator=g).cuda() for _ in range(4)]
    want = torch.mean(torch.stack(vols), dim=0)
    assert rel(vo.mean_fuse(vols).cpu(), want.cpu()) <= 1e-6
    assert torch.equal(vo.mean_fuse(vols[:1]), vols[0])


def test_target_pad_device_matches_numpy():
    from rehrseg_b200 import volume_ops as vo
    z = np.load(os.path.join(G, "volume_ops.npz"))
    padded, pads = vo.target_pad(torch.from_numpy(z["pad_in"]).cuda(), (10, 9, 7), mode="reflect")
    assert np.array_equal(padded.cpu().numpy(), z["pad_out"]) and np.array_equal(np.array(pads), z["pad_pads"])
    assert np.array_equal(vo.crop(padded, pads).cpu().numpy(), z["pad_crop"])


def test_postprocess_volumes_match_the_reference_arithmetic():
    """utils/sr_utils.py:244-304 after parse_image: zeroonenorm, z-first, F.conv2d(.., padding="same") blur along x, z back
    (`postprocess_flavr`), and the two blurred in-plane orientations of `postprocess_smore` -- vs the same lines in torch on the CPU."""
    import torch.nn.functional as F
    from rehrseg_b200 import volume_ops as vo
    g = torch.Generator().manual_seed(21)
    image = torch.randn((24, 20, 9), generator=g) * 3 + 1
    taps = torch.exp(-0.5 * ((torch.arange(9.) - 4) / 1.6) ** 2)
    k = (taps / taps.sum()).reshape(1, 1, 9, 1)
    x = image.numpy()
    x = (x - x.min()) / (x.max() - x.min()) * 255.0                                        # zeroonenorm
    want = F.conv2d(torch.from_numpy(x.transpose(2, 0, 1)).unsqueeze(1), k, padding="same").squeeze(1).numpy().transpose(1, 2, 0)
    got = vo.postprocess_flavr_volume(image.cuda(), k.cuda())
    assert tuple(got.shape) == want.shape
    assert float((got.cpu() - torch.from_numpy(want)).norm() / torch.from_numpy(want).norm()) < 1e-5
    vol = torch.stack([image, (image > 1).float()], dim=-1)                                # [X, Y, Z, 2]
    img_hr, label_hr, xr, yr = vo.postprocess_smore_volume(vol.cuda(), k.cuda())
    v = vol.numpy()
    want_x = F.conv2d(torch.from_numpy(v.transpose(2, 3, 0, 1))[:, 0:1], k, padding="same")
    want_y = F.conv2d(torch.from_numpy(v.transpose(2, 3, 1, 0).copy())[:, 0:1], k, padding="same")
    assert label_hr.dtype == torch.uint8 and tuple(img_hr.shape) == (24, 20, 9, 1)
    assert float((xr.cpu() - want_x).norm() / want_x.norm()) < 1e-5 and float((yr.cpu() - want_y).norm() / want_y.norm()) < 1e-5
