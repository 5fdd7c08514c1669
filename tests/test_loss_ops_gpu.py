"""GPU: the fused loss reductions (csrc/loss_ops.cu through rehrseg_b200/loss_ops.py) against the reference's own Distiller /
_build_loss outputs (tests/golden/joint_step.npz) and against the plain-PyTorch mirrors on ragged shapes.  fp32 bound of
north_star: relative L2 <= 1e-5 (summation order differs, nothing else)."""
import os

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu
G = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "joint_step.npz"))


def _t(name, grad=False):
    return torch.from_numpy(G[name]).cuda().requires_grad_(grad)


def rel(a, b):
    a, b = torch.as_tensor(a).double().cpu(), torch.as_tensor(b).double().cpu()
    return float((a - b).norm() / (b.norm() + 1e-30))


@pytest.mark.parametrize("tag,lam", [("cos_struct", (0.0, 1.0, 1.0)), ("all", (0.5, 1.0, 2.0))])
def test_fused_distiller_matches_reference_fixture(tag, lam):
    from rehrseg_b200 import loss_ops as lo
    d = lo.FusedDistiller(64, 64, *lam)
    assert list(d.state_dict().keys()) == list(G["distill_keys"])
    d.load_state_dict({"distill.weight": torch.from_numpy(G["distill_w"]), "distill.bias": torch.from_numpy(G["distill_b"])})
    d = d.cuda()
    fs, ft = _t("feat_s", True), _t("feat_t")
    with torch.backends.cudnn.flags(allow_tf32=False):
        loss = d(fs, ft)
        loss.backward()
    want = float(G[f"distill_{tag}_loss"])
    assert abs(float(loss.detach()) - want) <= 1e-5 * abs(want)
    assert rel(fs.grad, G[f"distill_{tag}_dfeat"]) < 2e-5
    assert rel(d.distill.weight.grad, G[f"distill_{tag}_dw"]) < 2e-5


def test_fused_seg_loss_matches_reference_fixture():
    from rehrseg_b200 import loss_ops as lo
    for tag, wd, unc in (("lr_unc", 0, _t("unc")), ("hr", 1, None), ("lr_nounc", 1, "omit")):
        obj = lo.build_fused_loss(False, weight_dice=wd)
        logits, target = _t("logits", True), _t("target")
        loss = obj(logits, target) if isinstance(unc, str) else obj(logits, target, unc)
        loss.backward()
        want = float(G[f"loss_{tag}"])
        assert abs(float(loss.detach()) - want) <= 1e-5 * max(1.0, abs(want)), tag
        assert rel(logits.grad, G[f"loss_{tag}_dlogits"]) < 2e-5, tag


@pytest.mark.parametrize("shape,classes", [((2, 5, 33, 17), 2), ((1, 3, 40, 70), 4), ((3, 1, 9, 1000), 3)])
def test_fused_seg_loss_vs_mirror_ragged(shape, classes):
    from rehrseg_b200 import loss_ops as lo, train_step as ts
    b, d, h, w = shape
    g = torch.Generator().manual_seed(3)
    logits = (3 * torch.randn((b, classes, d, h, w), generator=g)).cuda()
    target = torch.randint(0, classes, (b, 1, d, h, w), generator=g).float().cuda()
    unc = (torch.rand((b, 1, d, h, w), generator=g) * 0.99 + 0.01).cuda()
    for wd, u in ((0, unc), (1, None), (1, unc)):
        la, lb = logits.clone().requires_grad_(True), logits.clone().requires_grad_(True)
        fa = lo.FusedDCAndWeightedCELoss(1, wd)(la, target, u)
        fb = ts.DCAndWeightedCELoss(1, wd)(lb, target, u)
        fa.backward()
        fb.backward()
        assert abs(float(fa.detach()) - float(fb.detach())) <= 1e-5 * max(1.0, abs(float(fb.detach())))
        assert rel(la.grad, lb.grad) < 2e-5


@pytest.mark.parametrize("shape", [(2, 64, 3, 12, 10), (1, 16, 2, 9, 7), (2, 8, 4, 33, 65)])
def test_fused_distiller_pieces_vs_mirror_ragged(shape):
    """odd plane sizes exercise the ceil-mode partial windows of the pooled structure loss; C != 64 the generic channel loop"""
    from rehrseg_b200 import loss_ops as lo, train_step as ts
    g = torch.Generator().manual_seed(5)
    a = torch.randn(shape, generator=g).cuda()
    b = torch.randn(shape, generator=g).cuda()
    for fused, plain in ((lo.fused_cosine_distance_loss, ts.cosine_distance_loss), (lo.fused_structure_loss, ts.structure_loss)):
        x1, x2 = a.clone().requires_grad_(True), a.clone().requires_grad_(True)
        l1, l2 = fused(x1, b), plain(x2, b)
        l1.backward()
        l2.backward()
        assert abs(float(l1.detach()) - float(l2.detach())) <= 1e-5 * max(1e-3, abs(float(l2.detach()))), fused.__name__
        assert rel(x1.grad, x2.grad) < 2e-5, fused.__name__


def test_plane_maxpool_picks_the_first_maximum_like_aten():
    from rehrseg_b200 import loss_ops as lo
    x = torch.zeros((1, 2, 1, 6, 6), device="cuda")
    x[0, 0, 0, 1, 1] = x[0, 0, 0, 2, 0] = 5.0          # tie inside the first 3x3 window: ATen keeps the first in row-major order
    x.requires_grad_(True)
    xr = x.detach().clone().requires_grad_(True)
    out = lo.PlaneMaxPool.apply(x, 3, 3)
    ref = torch.nn.functional.max_pool2d(xr[:, :, 0], (3, 3), (3, 3), 0, ceil_mode=True)
    assert torch.equal(out, ref)
    w = torch.arange(out.numel(), dtype=torch.float32, device="cuda").reshape(out.shape) + 1
    (out * w).sum().backward()
    (ref * w).sum().backward()
    assert torch.equal(x.grad, xr.grad)


def test_loss_ops_refuse_cpu_tensors():
    from rehrseg_b200 import loss_ops as lo
    from rehrseg_b200._lib import RehrError
    with pytest.raises(RehrError):
        lo.fused_cosine_distance_loss(torch.randn(1, 4, 2, 4, 4), torch.randn(1, 4, 2, 4, 4))
