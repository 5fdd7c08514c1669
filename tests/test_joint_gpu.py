"""GPU: the stage-2 joint SR + segmentation step (BASELINE config 4; train_all.py:519-558) on the engine -- anisotropic
SegModel student (bf16 kernels through the C-ABI), UASR FLAVR teacher sweep, uncertainty-weighted CE + CE/Dice + structural
distillation -- against the oracle's fp32 CPU restatement of the same step with identical weights and batch."""
import pytest
import torch

pytestmark = pytest.mark.gpu


def rel(a, b):
    a, b = a.detach().double().cpu(), b.detach().double().cpu()
    return float((a - b).norm() / (b.norm() + 1e-30))


def _batch(seed, b, d, hw, up=4):
    g = torch.Generator().manual_seed(seed)
    img = torch.randn((b, 1, d, hw, hw), generator=g)
    label_lr = (torch.rand((b, 1, d, hw, hw), generator=g) > 0.8).float()
    label = (torch.rand((b, 1, d * up, hw, hw), generator=g) > 0.8).float()
    unc = torch.rand((b, 1, d, hw, hw), generator=g) * 0.99 + 0.01
    return img, label_lr, label, unc


@pytest.mark.parametrize("fused", [False, True])
def test_joint_step_vs_oracle(fused):
    """fused=True: Distiller / losses on the fused CUDA reductions (rehrseg_b200/loss_ops.py) instead of the plain-PyTorch mirrors"""
    from oracle import flavr as of, joint as oj, seg_model as ref_seg
    from rehrseg_b200 import flavr, loss_ops, seg_model as sm, train_step as ts

    ref_student = ref_seg.build("anisotropic")
    student = sm.SegModel(**ref_seg.plan_kwargs("anisotropic"))
    student.load_state_dict(ref_student.state_dict())
    student = student.cuda()
    ref_teacher = of.build(True).eval()
    torch.manual_seed(1234)
    teacher = flavr.UNet_3D_3D(2, "unet_18", 4, 4, False, "concat", "transpose", use_uncertainty=True).cuda().eval()
    torch.manual_seed(5)
    ref_dist = oj.RefDistiller(64, 64, 0.0, 1.0, 1.0)
    dist_mod = (loss_ops.FusedDistiller if fused else ts.Distiller)(64, 64, 0.0, 1.0, 1.0)
    dist_mod.load_state_dict(ref_dist.state_dict())
    dist_mod = dist_mod.cuda()

    batch = _batch(4, b=2, d=8, hw=64)
    want = oj.ref_joint_step(ref_student, tuple(t.clone() for t in batch), ref_teacher, ref_dist)
    gpu_batch = tuple(t.clone().cuda() for t in batch)
    build = loss_ops.build_fused_loss if fused else ts.build_loss
    got = ts.joint_train_step(student, gpu_batch, build(False, 0), build(False, 1), None, teacher, dist_mod)
    torch.cuda.synchronize()
    print({k: (float(got[k]), float(want[k])) for k in want})
    # scalar losses: means over >= 65k voxels of bf16-perturbed logits
    for k, tol in (("loss_lr_seg", 5e-3), ("loss_hr_seg", 5e-3), ("distill_loss", 2e-2), ("loss", 5e-3)):
        assert abs(float(got[k]) - float(want[k])) <= tol * max(1.0, abs(float(want[k]))), (k, float(got[k]), float(want[k]))
    # gradients through 22 InstanceNorm layers in bf16: same yardstick as tests/test_segmodel_gpu.py (global rel-L2 vs fp32)
    pr = dict(ref_student.named_parameters())
    num = den = 0.0
    seen = 0
    for name, p in student.named_parameters():
        if p.grad is None or pr[name].grad is None or (name.endswith("conv.bias") and ".convs." in name):
            continue
        a, b = p.grad.double().cpu(), pr[name].grad.double()
        num += float((a - b).pow(2).sum())
        den += float(b.pow(2).sum())
        seen += 1
    assert seen > 80
    assert (num / den) ** 0.5 <= 0.2, (num / den) ** 0.5
    # the distillation tap: the projection's gradient only flows through the distillation loss
    assert rel(dist_mod.distill.weight.grad, ref_dist.distill.weight.grad) <= 0.1
    # z-score mutated the caller's device image in place exactly like the reference (train_all.py:86)
    ref_img = batch[0].clone()
    from oracle import volume as ov
    ov.zscore_normalization(ref_img)
    assert rel(gpu_batch[0], ref_img) <= 1e-5


def test_joint_step_sgd_updates_parameters_and_caches():
    """Two consecutive steps with the reference's optimiser (train_all.py:513): the second step must see the updated weights
    (bf16 operand caches are keyed on the parameter version), i.e. its loss differs and matches a fresh model loaded with the
    updated state_dict."""
    from oracle import seg_model as ref_seg
    from rehrseg_b200 import seg_model as sm, train_step as ts
    torch.manual_seed(1234)
    student = sm.SegModel(**ref_seg.plan_kwargs("anisotropic")).cuda()
    opt = torch.optim.SGD(student.parameters(), lr=1e-2, momentum=0.99, nesterov=True, weight_decay=3e-5)
    batch = tuple(t.cuda() for t in _batch(9, b=1, d=8, hw=64))
    lr_obj, hr_obj = ts.build_loss(False, 0), ts.build_loss(False, 1)
    first = ts.joint_train_step(student, batch, lr_obj, hr_obj, opt)
    clone = sm.SegModel(**ref_seg.plan_kwargs("anisotropic")).cuda()
    clone.load_state_dict(student.state_dict())
    second = ts.joint_train_step(student, batch, lr_obj, hr_obj, opt)
    fresh = ts.joint_train_step(clone, batch, lr_obj, hr_obj, None)
    assert abs(float(second["loss"]) - float(first["loss"])) > 1e-5
    assert abs(float(second["loss"]) - float(fresh["loss"])) <= 1e-5 * max(1.0, abs(float(fresh["loss"])))


def test_graphed_joint_step_matches_the_python_launched_one():
    """train_step.GraphedJointStep (the whole stage-2 iteration replayed from one CUDA graph) against joint_train_step on the same
    weights and batch: same kernels, so losses and gradients agree to summation-order accuracy; a second replay with another batch
    must follow the new data (static inputs are re-loaded, the teacher's in-place z-score included)."""
    from rehrseg_b200 import flavr, loss_ops, seg_model as sm, train_step as ts
    from oracle import seg_model as ref_seg
    torch.manual_seed(7)
    student = sm.SegModel(**ref_seg.plan_kwargs("anisotropic")).cuda()
    teacher = flavr.UNet_3D_3D(2, "unet_18", 4, 4, False, "concat", "transpose", use_uncertainty=True).cuda().eval()
    dist_mod = loss_ops.FusedDistiller(64, 64, 0.0, 1.0, 1.0).cuda()
    lr_obj, hr_obj = loss_ops.build_fused_loss(False, 0), loss_ops.build_fused_loss(False, 1)
    params = list(student.parameters()) + list(dist_mod.parameters())
    batches = [tuple(t.cuda() for t in _batch(s, b=2, d=8, hw=64)) for s in (4, 5)]

    want = []
    for b in batches:
        out = ts.joint_train_step(student, tuple(t.clone() for t in b), lr_obj, hr_obj, None, teacher, dist_mod)
        torch.cuda.synchronize()
        want.append(({k: float(v) for k, v in out.items()}, [None if p.grad is None else p.grad.clone() for p in params]))

    step = ts.GraphedJointStep(student, tuple(t.clone() for t in batches[0]), lr_obj, hr_obj, teacher, dist_mod)
    for b, (terms, grads) in zip(batches, want):
        got = step(tuple(t.clone() for t in b))
        torch.cuda.synchronize()
        for k, v in terms.items():
            assert abs(float(got[k]) - v) <= 1e-4 * max(1.0, abs(v)), (k, float(got[k]), v)
        for p, g in zip(params, grads):
            assert (p.grad is None) == (g is None)
            if g is not None:
                assert rel(p.grad, g) <= 2e-3, rel(p.grad, g)
    step.close()
