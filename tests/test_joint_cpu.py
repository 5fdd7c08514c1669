"""CPU: the stage-2 joint-step pieces (BASELINE config 4).
  * oracle/joint.py (restatement) against tests/golden/joint_step.npz, generated from the reference's own Distiller and
    _build_loss (oracle/make_golden.py) -- this pins the oracle;
  * the host-side mirrors in rehrseg_b200/train_step.py (plain PyTorch, no kernels) against the same fixture;
  * gloo world-size-2: `joint_train_step` + `allreduce_gradients` give every rank the mean of the two ranks' gradients
    (the networks on this path are CPU stand-ins: the oracle modules -- the engine modules need a GPU)."""
import os
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from oracle import joint as oj
from rehrseg_b200 import train_step as ts

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
G = np.load(os.path.join(ROOT, "tests", "golden", "joint_step.npz"))


def _t(name, grad=False):
    return torch.from_numpy(G[name]).clone().requires_grad_(grad)


def _rel(a, b):
    a, b = torch.as_tensor(a).double(), torch.as_tensor(b).double()
    return float((a - b).norm() / (b.norm() + 1e-30))


@pytest.mark.parametrize("cls", [oj.RefDistiller, ts.Distiller])
@pytest.mark.parametrize("tag,lam", [("cos_struct", (0.0, 1.0, 1.0)), ("all", (0.5, 1.0, 2.0))])
def test_distiller_matches_reference_fixture(cls, tag, lam):
    d = cls(64, 64, *lam)
    assert list(d.state_dict().keys()) == list(G["distill_keys"])
    d.load_state_dict({"distill.weight": _t("distill_w"), "distill.bias": _t("distill_b")})
    fs, ft = _t("feat_s", True), _t("feat_t")
    loss = d(fs, ft)
    loss.backward()
    assert abs(float(loss) - float(G[f"distill_{tag}_loss"])) <= 1e-6 * abs(float(G[f"distill_{tag}_loss"]))
    assert _rel(fs.grad, G[f"distill_{tag}_dfeat"]) < 1e-5
    assert _rel(d.distill.weight.grad, G[f"distill_{tag}_dw"]) < 1e-5


@pytest.mark.parametrize("build", [oj.ref_build_loss, ts.build_loss])
def test_seg_loss_matches_reference_fixture(build):
    for tag, wd, unc in (("lr_unc", 0, _t("unc")), ("hr", 1, None), ("lr_nounc", 1, "omit")):
        obj = build(False, weight_dice=wd)
        logits, target = _t("logits", True), _t("target")
        loss = obj(logits, target) if isinstance(unc, str) else obj(logits, target, unc)
        loss.backward()
        assert abs(float(loss) - float(G[f"loss_{tag}"])) <= 2e-6 * max(1.0, abs(float(G[f"loss_{tag}"]))), tag
        assert _rel(logits.grad, G[f"loss_{tag}_dlogits"]) < 1e-5, tag


def test_uncertainty_broadcast_quirk_is_kept():
    """utils/seg_utils.py:299-301,349: CE [B,D,H,W] x uncertainty [B,1,D,H,W] broadcasts over the batch -- the mirror must
    weight every sample's CE by every sample's uncertainty, not pair them."""
    logits, target, unc = _t("logits"), _t("target"), _t("unc")
    ce = torch.nn.functional.cross_entropy(logits, target[:, 0].long(), reduction="none")
    paired = (ce * unc[:, 0]).mean()
    crossed = (ce[None] * unc).mean()
    got = ts.build_loss(False, 0)(logits, target, unc)
    assert abs(float(got) - float(crossed)) < 1e-6 and abs(float(got) - float(paired)) > 1e-4


def _batch(seed, b=1, d=5, hw=32, up=4):
    g = torch.Generator().manual_seed(seed)
    img = torch.randn((b, 1, d, hw, hw), generator=g)
    label_lr = (torch.rand((b, 1, d, hw, hw), generator=g) > 0.8).float()
    label = (torch.rand((b, 1, d * up, hw, hw), generator=g) > 0.8).float()
    unc = torch.rand((b, 1, d, hw, hw), generator=g) * 0.99 + 0.01
    return img, label_lr, label, unc


def _tiny_student():
    from oracle import seg_model as ref_seg
    kw = ref_seg.plan_kwargs("tiny")
    kw.update(kernel_sizes=[[1, 3, 3], [3, 3, 3], [3, 3, 3]], strides=[[1, 1, 1], [1, 2, 2], [1, 2, 2]], features_per_stage=[32, 64, 128])
    torch.manual_seed(1234)
    return ref_seg.RefSegModel(**kw)


def test_joint_step_mirror_equals_oracle_step_on_cpu_standins():
    """Same CPU networks through the product's `joint_train_step` and the oracle's `ref_joint_step`: identical losses and
    gradients (the step logic -- in-place z-score shared by teacher and student, loss wiring, distillation tap -- is the
    only thing under test here)."""
    from oracle import flavr as ref_flavr
    teacher = ref_flavr.build(use_uncertainty=True).eval()
    res = []
    for which in ("mirror", "oracle"):
        student = _tiny_student()
        torch.manual_seed(5)
        dist_mod = (ts.Distiller if which == "mirror" else oj.RefDistiller)(64, 64, 0.0, 1.0, 1.0)
        batch = tuple(t.clone() for t in _batch(4))
        if which == "mirror":
            out = ts.joint_train_step(student, batch, ts.build_loss(False, 0), ts.build_loss(False, 1), None, teacher, dist_mod,
                                      device=torch.device("cpu"))
        else:
            out = oj.ref_joint_step(student, batch, teacher, dist_mod)
        grads = torch.cat([p.grad.reshape(-1) for p in student.parameters() if p.grad is not None])
        res.append((out, grads, batch[0]))
    (a, ga, ia), (b, gb, ib) = res
    for k in ("loss", "loss_lr_seg", "loss_hr_seg", "distill_loss"):
        assert abs(float(a[k]) - float(b[k])) <= 1e-6 * max(1.0, abs(float(b[k]))), k
    assert _rel(ga, gb) < 1e-5
    assert torch.equal(ia, ib) and abs(float(ia.mean())) < 1e-5      # the caller's image was z-scored in place


def _dp_worker(rank, world, port, q):
    sys.path.insert(0, ROOT)
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    torch.set_num_threads(2)
    from rehrseg_b200 import train_step as ts2
    student = _tiny_student()
    batch = _batch(40 + rank, d=4, hw=16)
    out = ts2.joint_train_step(student, batch, ts2.build_loss(False, 0), ts2.build_loss(False, 1), None, device=torch.device("cpu"))
    flat = torch.cat([p.grad.reshape(-1) for p in student.parameters() if p.grad is not None])
    q.put((rank, float(out["loss"]), flat.numpy()))
    dist.barrier()
    dist.destroy_process_group()


def test_joint_step_data_parallel_world2():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 31500 + (os.getpid() % 2000)
    procs = [ctx.Process(target=_dp_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    got = sorted([q.get(timeout=300) for _ in procs], key=lambda t: t[0])
    for p in procs:
        p.join(60)
        assert p.exitcode == 0
    assert np.array_equal(got[0][2], got[1][2])            # both ranks hold the same averaged gradient
    # single-process check: mean of the two ranks' local gradients
    local = []
    for rank in range(2):
        student = _tiny_student()
        oj.ref_joint_step(student, _batch(40 + rank, d=4, hw=16))
        local.append(torch.cat([p.grad.reshape(-1) for p in student.parameters() if p.grad is not None]))
    assert _rel(got[0][2], (local[0] + local[1]) / 2) < 1e-5


def test_oracle_joint_step_reproduces_the_reference_loop_body():
    """tests/golden/joint_whole_step.npz was produced by the reference's own SegModel / UNet_3D_3D / Distiller / _build_loss /
    get_intermediate_features executing train_all.py:520-555 (oracle/make_golden.py:joint_step_fixture).  The oracle's modules have
    the same default initialisation, so `ref_joint_step` must reproduce every loss term and gradient -- this pins the restated step."""
    from oracle import flavr as ref_flavr
    from oracle import seg_model as ref_seg
    W = np.load(os.path.join(ROOT, "tests", "golden", "joint_whole_step.npz"))
    kw = ref_seg.plan_kwargs("tiny")
    kw.update(n_stages=3, features_per_stage=[32, 64, 128], kernel_sizes=[[1, 3, 3], [3, 3, 3], [3, 3, 3]],
              strides=[[1, 1, 1], [1, 2, 2], [1, 2, 2]], n_conv_per_stage=[2] * 3, n_conv_per_stage_decoder=[2] * 2)
    torch.manual_seed(1234)
    student = ref_seg.RefSegModel(**kw)
    teacher = ref_flavr.build(use_uncertainty=True, seed=1234).eval()
    torch.manual_seed(5)
    dist_mod = oj.RefDistiller(64, 64, 0.0, 1.0, 1.0)
    batch = tuple(torch.from_numpy(W[k]).clone() for k in ("img", "label_lr", "label", "uncertainty_lr"))
    out = oj.ref_joint_step(student, batch, teacher, dist_mod)
    for k in ("loss", "loss_lr_seg", "loss_hr_seg", "distill_loss"):
        assert abs(float(out[k]) - float(W[k])) <= 2e-6 * max(1.0, abs(float(W[k]))), (k, float(out[k]), float(W[k]))
    assert _rel(batch[0], W["img_after"]) < 1e-6                       # in-place z-score of the caller's image
    assert _rel(student.encoder.stages[0][0].convs[0].conv.weight.grad, W["grad_stem"]) < 1e-4
    assert _rel(student.sr_head[0].weight.grad, W["grad_sr_head0"]) < 1e-4
    assert _rel(dist_mod.distill.weight.grad, W["grad_distill"]) < 1e-4
    tot = float(sum(p.grad.double().abs().sum() for p in student.parameters() if p.grad is not None))
    assert abs(tot - float(W["grad_abs_sum"])) <= 1e-4 * float(W["grad_abs_sum"])
    # the product-side mirror of the step on the same CPU stand-ins gives the same numbers
    torch.manual_seed(1234)
    student2 = ref_seg.RefSegModel(**kw)
    torch.manual_seed(5)
    dist2 = ts.Distiller(64, 64, 0.0, 1.0, 1.0)
    batch2 = tuple(torch.from_numpy(W[k]).clone() for k in ("img", "label_lr", "label", "uncertainty_lr"))
    got = ts.joint_train_step(student2, batch2, ts.build_loss(False, 0), ts.build_loss(False, 1), None, teacher, dist2,
                              device=torch.device("cpu"))
    for k in ("loss", "loss_lr_seg", "loss_hr_seg", "distill_loss"):
        assert abs(float(got[k]) - float(W[k])) <= 2e-6 * max(1.0, abs(float(W[k]))), k


# ---------------------------------------------------------------------------------------------------------------
# SR stage: the loop body of train_sr (train_all.py:118-139) incl. the UASR terms, pinned by tests/golden/sr_step.npz, which the
# reference's OWN train_sr produced on its own UNet_3D_3D / BCEDiceLoss (oracle/make_golden.py::sr_step_fixture)
# ---------------------------------------------------------------------------------------------------------------
S = np.load(os.path.join(ROOT, "tests", "golden", "sr_step.npz"))


@pytest.mark.parametrize("unc", [False, True])
def test_oracle_sr_step_reproduces_the_reference_train_sr(unc):
    from oracle import flavr as of
    tag = "uasr" if unc else "plain"
    net = of.build(unc, seed=1234)
    lr, hr = torch.from_numpy(S["patches_lr"]).clone(), torch.from_numpy(S["patches_hr"]).clone()
    loss = oj.ref_sr_step(net, lr, hr, torch.nn.L1Loss(), oj.RefBCEDiceLoss(1, 1), 4, 4, unc)
    assert abs(float(loss) - float(S[f"{tag}_loss"])) <= 2e-6 * abs(float(S[f"{tag}_loss"]))
    named = dict(net.named_parameters())
    assert _rel(named["encoder.stem.0.weight"].grad, S[f"{tag}_grad_stem"]) < 1e-4
    head = named["uncertainty_out.weight"] if unc else named["outconv.1.weight"]
    assert _rel(head.grad, S[f"{tag}_grad_outconv"]) < 1e-4
    tot = float(sum(p.grad.double().abs().sum() for p in net.parameters() if p.grad is not None))
    assert abs(tot - float(S[f"{tag}_grad_abs_sum"])) <= 1e-4 * float(S[f"{tag}_grad_abs_sum"])


@pytest.mark.parametrize("unc", [False, True])
def test_sr_train_step_mirror_matches_the_reference_fixture(unc):
    """rehrseg_b200.train_step.sr_train_step + BCEDiceLoss (host logic, plain PyTorch) around a CPU stand-in network (the oracle
    FLAVR: the engine module needs a GPU, tests/test_flavr_gpu.py drives the same function through it)."""
    from oracle import flavr as of
    tag = "uasr" if unc else "plain"
    net = of.build(unc, seed=1234)
    lr, hr = torch.from_numpy(S["patches_lr"]).clone(), torch.from_numpy(S["patches_hr"]).clone()
    opt = torch.optim.SGD(net.parameters(), lr=0.0)
    out = ts.sr_train_step(net, (lr, hr), torch.nn.L1Loss(), ts.BCEDiceLoss(1, 1), opt, None, 4, 4, unc, device=torch.device("cpu"))
    assert abs(float(out["loss"]) - float(S[f"{tag}_loss"])) <= 2e-6 * abs(float(S[f"{tag}_loss"]))
    assert _rel(dict(net.named_parameters())["encoder.stem.0.weight"].grad, S[f"{tag}_grad_stem"]) < 1e-4
    x = torch.randn(3, 2, 4, 8, 8)
    t = (torch.rand(3, 2, 4, 8, 8) > 0.5).float()
    assert abs(float(ts.BCEDiceLoss(0.3, 0.7)(x, t)) - float(oj.RefBCEDiceLoss(0.3, 0.7)(x, t))) < 1e-6
