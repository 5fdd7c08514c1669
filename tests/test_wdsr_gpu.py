"""GPU: WDSR (SMORE 2-D stage, models/wdsr.py) on the engine vs the oracle restatement (fp32 CPU) and the fixture the reference's
own module produced (tests/golden/wdsr_small.npz, `resize` by the factor 1 = identity).  bf16 bound: relative L2 <= 1e-2."""
import os
import warnings

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu
G = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def rel(a, b):
    a, b = a.detach().double().cpu(), b.detach().double().cpu()
    return float((a - b).norm() / (b.norm() + 1e-30))


def _pair(seed=1234, **kw):
    from oracle import wdsr as ow
    from rehrseg_b200 import wdsr
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        ref = ow.build(seed=seed, **kw)
        torch.manual_seed(seed)
        mine = wdsr.WDSR(kw.get("out_channel", 2), kw.get("n_resblocks", 2), kw.get("num_channels", 32), kw.get("scale", 4.0))
    assert list(ref.state_dict().keys()) == list(mine.state_dict().keys())
    assert all(torch.equal(a, b) for a, b in zip(ref.state_dict().values(), mine.state_dict().values()))   # same default init
    return ref, mine.cuda()


def test_wdsr_forward_vs_reference_fixture():
    z = np.load(os.path.join(G, "wdsr_small.npz"))
    ref, mine = _pair()
    assert list(mine.state_dict().keys()) == list(z["keys"])
    out = mine(torch.from_numpy(z["x"]).cuda())
    assert tuple(out.shape) == z["out"].shape and out.dtype == torch.float32
    assert rel(out, torch.from_numpy(z["out"])) <= 1e-2
    assert mine.calc_out_patch_size([16, 12]) == list(z["patch"])


def test_wdsr_fwd_bwd_vs_oracle_full_depth():
    """The configuration train_all.py:267-272 builds (out_channel 2, 16 blocks, 32 channels, scale 4) on a ragged patch."""
    ref, mine = _pair(seed=5, n_resblocks=16)
    g = torch.Generator().manual_seed(2)
    x = torch.rand((2, 2, 40, 36), generator=g)
    out_r = ref(x)
    out_m = mine(x.cuda())
    assert rel(out_m, out_r) <= 1e-2
    cot = torch.randn(out_r.shape, generator=g)
    (out_r * cot).sum().backward()
    (out_m * cot.cuda()).sum().backward()
    pr = dict(ref.named_parameters())
    num = den = 0.0
    for name, p in mine.named_parameters():
        assert p.grad is not None, name
        a, b = p.grad.double().cpu(), pr[name].grad.double()
        num += float((a - b).pow(2).sum()); den += float(b.pow(2).sum())
    print("wdsr grads global", (num / den) ** 0.5)
    assert (num / den) ** 0.5 <= 5e-2


def test_wdsr_rational_scale_is_refused_and_smore_sweep():
    from rehrseg_b200 import wdsr
    from rehrseg_b200._lib import RehrError
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        m = wdsr.WDSR(2, 1, 32, 2.5).cuda()
    with pytest.raises(RehrError):
        m(torch.rand((1, 2, 8, 8), device="cuda"))
    ref, mine = _pair()
    vol = torch.rand((5, 2, 12, 16), generator=torch.Generator().manual_seed(4))          # [slices, C, y, x]
    got = wdsr.apply_to_vol_smore(mine, vol.cuda(), 2)
    with torch.no_grad():
        want = torch.cat([ref(vol[i:i + 2].permute(0, 1, 3, 2)) for i in range(0, 5, 2)], 0)
    assert got.shape == want.shape and not got.is_cuda and rel(got, want) <= 1e-2
