"""GPU: FLAVR UNet_3D_3D on the engine (bf16, through the C-ABI) vs the oracle restatement (fp32 CPU) and the fixtures the
reference's own module produced.  bf16 bound of north_star: relative L2 <= 1e-2 on the SR volumes."""
import os

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu
G = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def rel(a, b):
    a, b = a.detach().double().cpu(), b.detach().double().cpu()
    return float((a - b).norm() / (b.norm() + 1e-30))


def _pair(unc, seed=1234):
    from oracle import flavr as of
    from rehrseg_b200 import flavr
    ref = of.build(unc, seed=seed).eval()
    torch.manual_seed(seed)
    mine = flavr.UNet_3D_3D(2, "unet_18", 4, 4, False, "concat", "transpose", use_uncertainty=unc)
    assert all(torch.equal(a, b) for a, b in zip(ref.state_dict().values(), mine.state_dict().values()))  # same default init
    return ref, mine.cuda().eval()


@pytest.mark.parametrize("unc", [False, True])
def test_flavr_forward_vs_reference_fixture(unc):
    z = np.load(os.path.join(G, "flavr_small.npz"))
    _, mine = _pair(unc)
    x = torch.from_numpy(z["x"]).cuda()
    tag = "uasr" if unc else "plain"
    with torch.no_grad():
        xin = x.clone()
        out = mine(xin)
    assert rel(xin, torch.from_numpy(z[f"{tag}_x_after"])) <= 1e-6      # same in-place mutation of the caller's tensor
    if unc:
        assert out[0].shape == z["uasr_out"].shape and out[1].shape == z["uasr_unc"].shape
        assert rel(out[0], torch.from_numpy(z["uasr_out"])) <= 1e-2
        assert rel(out[1], torch.from_numpy(z["uasr_unc"])) <= 1e-2
    else:
        assert out.shape == z["plain_out"].shape
        assert rel(out, torch.from_numpy(z["plain_out"])) <= 1e-2
        with torch.no_grad():
            f = mine(x.clone(), return_inetermediate_feature=True)
        assert len(f) == 5 and f[1].dtype == torch.float32
        assert rel(f[1], torch.from_numpy(z["plain_x1"])) <= 1e-2 and rel(f[4], torch.from_numpy(z["plain_x4"])) <= 1e-2


@pytest.mark.parametrize("unc", [False, True])
def test_flavr_fwd_bwd_vs_oracle(unc):
    ref, mine = _pair(unc, seed=7)
    mine.train()
    g = torch.Generator().manual_seed(3)
    x = torch.rand((2, 2, 4, 64, 48), generator=g)
    x[:, 1] = (x[:, 1] > 0.8).float()
    out_r = ref(x.clone())
    out_m = mine(x.clone().cuda())
    outs_r = out_r if unc else (out_r,)
    outs_m = out_m if unc else (out_m,)
    loss_r = loss_m = 0
    for i, (a, b) in enumerate(zip(outs_m, outs_r)):
        assert a.shape == b.shape and rel(a, b) <= 1e-2, (i, rel(a, b))
        cot = torch.randn(b.shape, generator=g)
        loss_r = loss_r + (b * cot).sum() / b.numel()
        loss_m = loss_m + (a * cot.cuda()).sum() / a.numel()
    loss_r.backward()
    loss_m.backward()
    pr = dict(ref.named_parameters())
    num = den = 0.0
    worst = (0.0, "")
    for name, p in mine.named_parameters():
        assert (p.grad is None) == (pr[name].grad is None), name
        if p.grad is None:
            continue
        a, b = p.grad.double().cpu(), pr[name].grad.double()
        num += float((a - b).pow(2).sum()); den += float(b.pow(2).sum())
        r = float((a - b).norm() / (b.norm() + 1e-30))
        if r > worst[0]:
            worst = (r, name)
    print("grads global", (num / den) ** 0.5, "worst", worst)
    assert (num / den) ** 0.5 <= 5e-2, ((num / den) ** 0.5, worst)


def test_get_intermediate_features_and_window_sweep():
    from oracle import flavr as of, volume as ov
    from rehrseg_b200 import flavr
    z = np.load(os.path.join(G, "flavr_small.npz"))
    _, teacher = _pair(True)
    img = torch.from_numpy(z["gif_img"]).cuda()
    lab = torch.from_numpy(z["gif_lab"]).cuda()
    with torch.no_grad():
        feats = flavr.get_intermediate_features(teacher, img, lab, normalize=flavr.zscore_normalization, max_batch=3)
    assert rel(img, torch.from_numpy(z["gif_img_after"])) <= 1e-6      # z-score mutated the caller's tensor, as in the reference
    assert sorted(feats.keys()) == [0, 1, 2, 3, 4]
    for i, key in ((1, "gif_f1"), (3, "gif_f3")):
        assert feats[i].shape == z[key].shape and rel(feats[i], torch.from_numpy(z[key])) <= 1e-2, key
    # opt-in `keys` (the stage-2 loop reads only [1]): the teacher stops after layer1, the returned map is the same tensor bit for bit
    with torch.no_grad():
        only1 = flavr.get_intermediate_features(teacher, torch.from_numpy(z["gif_img"]).cuda(), lab, normalize=flavr.zscore_normalization,
                                                max_batch=3, keys=(1,))
    assert sorted(only1.keys()) == [1] and torch.equal(only1[1], feats[1])
    # apply_to_vol_flavr: batched sweep vs the oracle's one-window-at-a-time restatement (ragged in-plane size -> pad to 16)
    ref, mine = _pair(False, seed=9)
    vol = torch.rand((6, 2, 40, 24), generator=torch.Generator().manual_seed(6))
    want = ov.apply_to_vol_flavr(ref, vol.clone())
    got = flavr.apply_to_vol_flavr(mine, vol.clone().cuda(), max_batch=2)
    assert got.shape == want.shape == (20, 2, 24, 40)
    assert rel(got, want) <= 1e-2


def test_window_sweep_vs_reference_function_fixture():
    """Engine sweep against the output of the reference's OWN apply_to_vol_flavr on its own network (tests/golden/sr_sweep.npz,
    oracle/make_golden.py:sr_sweep_fixture): same default-initialised weights, ragged 20 x 24 planes, 4 windows."""
    from rehrseg_b200 import flavr
    z = np.load(os.path.join(G, "sr_sweep.npz"))
    _, mine = _pair(False, seed=1234)
    got = flavr.apply_to_vol_flavr(mine, torch.from_numpy(z["vol"]).clone().cuda(), max_batch=3)
    want = torch.from_numpy(z["out"])
    assert got.shape == want.shape
    assert rel(got, want) <= 1e-2


def test_convert_reference_shaped_flavr():
    from oracle import flavr as of
    from rehrseg_b200 import flavr
    ref = of.build(False, seed=11).eval()
    x = torch.rand((1, 2, 4, 32, 32), generator=torch.Generator().manual_seed(1))
    with torch.no_grad():
        want = ref(x.clone())
    ids = {k: id(p) for k, p in ref.named_parameters()}
    m = flavr.convert(ref).cuda()
    assert {k: id(p) for k, p in m.named_parameters()} == ids
    with torch.no_grad():
        got = m(x.clone().cuda())
    assert rel(got, want) <= 1e-2
    assert m.calc_out_patch_size([4, 32, 32]) == [16, 32, 32] if hasattr(m, "calc_out_patch_size") else True


def test_sr_volume_orientations_vs_oracle():
    """Tensor part of inference_flavr (utils/sr_utils.py:157-175) with two orientations (0 and 180 degrees: the reference un-rotates over the OUTPUT dims (0, 1), so only shape-preserving angles can be stacked), mean fusion."""
    from oracle import volume as ov
    from rehrseg_b200 import flavr
    ref, mine = _pair(False, seed=13)
    img = torch.rand((7, 7, 20, 2), generator=torch.Generator().manual_seed(8))    # (hr, hr, lr, C), square in-plane
    with torch.no_grad():
        want = ov.sr_volume_orientations(ref, img.clone(), angles=(0, 180))
        got = flavr.sr_volume_orientations(mine, img.clone().cuda(), angles=(0, 180), max_batch=3)
    assert got.shape == want.shape
    assert rel(got, want) <= 1e-2


@pytest.mark.parametrize("unc", [False, True])
def test_sr_train_step_on_the_engine_vs_oracle(unc):
    """BASELINE config 2's training step, `train_sr` (train_all.py:118-139) with the UASR loss terms (:125-130): the engine FLAVR
    driven by rehrseg_b200.train_step.sr_train_step vs the oracle step on the fp32 CPU model -- loss and gradients."""
    from oracle import joint as oj
    from rehrseg_b200 import train_step as ts
    ref, mine = _pair(unc, seed=9)
    ref.train()
    mine.train()
    g = torch.Generator().manual_seed(13)
    lr = torch.rand((2, 2, 4, 64, 64), generator=g)
    lr[:, 1] = (lr[:, 1] > 0.8).float()
    hr = torch.rand((2, 2, 16, 64, 64), generator=g)
    hr[:, 1] = (hr[:, 1] > 0.8).float()
    want = float(oj.ref_sr_step(ref, lr.clone(), hr.clone(), torch.nn.L1Loss(), oj.RefBCEDiceLoss(1, 1), 4, 4, unc))
    lr_dev = lr.clone().cuda()
    got = float(ts.sr_train_step(mine, (lr_dev, hr.clone()), torch.nn.L1Loss(), ts.BCEDiceLoss(1, 1), None, None, 4, 4, unc)["loss"])
    assert abs(got - want) <= 1e-2 * abs(want), (got, want)
    # the reference's forward subtracts the channel-0 mean from the caller's batch in place (FLAVR_arch.py:180-181)
    assert rel(lr_dev[:, 0], lr[:, 0] - lr[:, 0:1].mean((2, 3, 4), keepdim=True)[:, 0]) <= 1e-5
    pr = dict(ref.named_parameters())
    num = den = 0.0
    for name, p in mine.named_parameters():
        assert (p.grad is None) == (pr[name].grad is None), name
        if p.grad is None:
            continue
        a, b = p.grad.double().cpu(), pr[name].grad.double()
        num += float((a - b).pow(2).sum()); den += float(b.pow(2).sum())
    print("sr step", "uasr" if unc else "plain", "loss", got, want, "grads global", (num / den) ** 0.5)
    assert (num / den) ** 0.5 <= 8e-2


def test_fused_uasr_mixture_matches_the_pytorch_formulation(monkeypatch):
    """The UASR head's expert mixture as one kernel per direction (flavr.uasr_mixture, rehr_uasr_mixture_fwd / _bwd) against the
    PyTorch statements of FLAVR_arch.py:203-227,244-246 on the same network: outputs, and the gradients of every parameter."""
    from rehrseg_b200 import flavr
    torch.manual_seed(3)
    net = flavr.UNet_3D_3D(2, "unet_18", 4, 4, False, "concat", "transpose", use_uncertainty=True).cuda()
    x = torch.rand((2, 2, 4, 64, 48), device="cuda")
    g = torch.Generator(device="cuda").manual_seed(4)

    def run(fused):
        monkeypatch.setattr(flavr, "FUSED_UASR", fused)
        for p in net.parameters():
            p.grad = None
        res, unc = net(x.clone())
        cot_r = torch.randn(res.shape, device="cuda", generator=torch.Generator(device="cuda").manual_seed(5))
        cot_u = torch.randn(unc.shape, device="cuda", generator=torch.Generator(device="cuda").manual_seed(6))
        ((res * cot_r).mean() + (unc * cot_u).mean()).backward()
        torch.cuda.synchronize()
        return res.detach(), unc.detach(), {n: p.grad.clone() for n, p in net.named_parameters() if p.grad is not None}

    r1, u1, g1 = run(True)
    r0, u0, g0 = run(False)
    assert r1.shape == r0.shape == (2, 2, 4, 64, 48) and u1.shape == u0.shape == (2, 1, 4, 64, 48)
    assert float((r1 - r0).abs().max()) <= 2e-6 * max(1.0, float(r0.abs().max()))
    assert float((u1 - u0).abs().max()) <= 2e-6
    assert set(g1) == set(g0)
    for n in g0:
        d = float((g1[n] - g0[n]).norm() / (g0[n].norm() + 1e-30))
        # the head's own layers see fp32 gradients on both paths; everything upstream receives them through one bf16 rounding
        assert d <= (1e-4 if n.startswith("uncertainty_out") else 2e-2), (n, d)
