"""CPU restatement (plain PyTorch fp32) of the stage-2 joint training step (BASELINE config 4) -- TEST INFRASTRUCTURE.

Follows, statement by statement:
  * RefDistiller / ref_cosine_distance / ref_similarity / ref_sim_dis / ref_pairwise_after_pool
        <- models/seg_model.py:60-151 (Distiller, cosine_distance_loss, L2, similarity, sim_dis_compute,
           CriterionPairWiseforWholeFeatAfterPool)
  * RefRobustCE / RefDCAndWeightedCE / ref_build_loss
        <- utils/seg_utils.py:289-372 (RobustCrossEntropyLoss, DC_and_weighted_CE_loss, _build_loss); the Dice class is the
           third-party nnunetv2 2.3.1 MemoryEfficientSoftDiceLoss restated in oracle/third_party.py (parity unpinned)
  * ref_joint_step
        <- train_all.py:519-558 (loop body of the stage-2 training) with the teacher sweep train_all.py:85-112
           (oracle/flavr.py:intermediate_features) and zscore_normalization utils/seg_utils.py:137-148 (oracle/volume.py)
Pinned against the reference's own Distiller and _build_loss by tests/golden/joint_step.npz (oracle/make_golden.py) and,
when /root/reference exists, live in tests/test_oracle_vs_reference.py.
"""
from __future__ import annotations

import numpy as np
import torch
import torch.nn.functional as F
from torch import nn

from . import flavr as ref_flavr
from . import third_party as tp
from . import volume as ref_vol


# ---- models/seg_model.py:60-78 ---------------------------------------------------------------------------------
def ref_cosine_distance(t1, t2):
    t1 = F.normalize(t1, p=2, dim=1)
    t2 = F.normalize(t2, p=2, dim=1)
    t1 = t1.reshape(t1.shape[0], t1.shape[1], -1)
    t2 = t2.reshape(t2.shape[0], t2.shape[1], -1)
    return (1 - torch.cosine_similarity(t1, t2, dim=2)).mean()


# ---- models/seg_model.py:80-93 ---------------------------------------------------------------------------------
def ref_similarity(feat):
    feat = feat.float()
    l2 = (((feat ** 2).sum(dim=1)) ** 0.5).reshape(feat.shape[0], 1, feat.shape[2], feat.shape[3]) + 1e-8
    feat = feat / l2.detach()
    feat = feat.reshape(feat.shape[0], feat.shape[1], -1)
    return torch.einsum("icm,icn->imn", [feat, feat])


def ref_sim_dis(f_s, f_t):
    err = ((ref_similarity(f_t) - ref_similarity(f_s)) ** 2) / ((f_t.shape[-1] * f_t.shape[-2]) ** 2) / f_t.shape[0]
    return err.sum()


# ---- models/seg_model.py:95-113 --------------------------------------------------------------------------------
def ref_pairwise_after_pool(preds_s, preds_t, scale):
    _, _, s, total_w, total_h = preds_s.shape
    fs = preds_s.permute(0, 2, 1, 3, 4).reshape(-1, preds_s.shape[1], total_w, total_h)      # 'b c s h w -> (b s) c h w'
    ft = preds_t.permute(0, 2, 1, 3, 4).reshape(-1, preds_t.shape[1], total_w, total_h)
    pw, ph = int(total_w * scale), int(total_h * scale)
    pool = nn.MaxPool2d(kernel_size=(pw, ph), stride=(pw, ph), padding=0, ceil_mode=True)
    return ref_sim_dis(pool(fs), pool(ft)) / s


# ---- models/seg_model.py:115-151 -------------------------------------------------------------------------------
class RefDistiller(nn.Module):
    def __init__(self, student_dim, teacher_dim, lambda_l1=0.0, lambda_cosine=0.0, lambda_structure=0.0):
        super().__init__()
        self.lambda_l1, self.lambda_cosine, self.lambda_structure = lambda_l1, lambda_cosine, lambda_structure
        self.distill = nn.Conv3d(in_channels=student_dim, out_channels=teacher_dim, kernel_size=1, stride=1, padding=0)

    def forward(self, feature_student, feature_teacher):
        loss = 0
        if self.lambda_structure > 0:
            loss += self.lambda_structure * ref_pairwise_after_pool(feature_student, feature_teacher, 0.5)
        distilled = self.distill(feature_student)
        if self.lambda_l1 > 0:
            loss += F.smooth_l1_loss(distilled, feature_teacher) * self.lambda_l1
        if self.lambda_cosine > 0:
            loss += self.lambda_cosine * ref_cosine_distance(distilled, feature_teacher)
        return loss


# ---- utils/seg_utils.py:289-303 --------------------------------------------------------------------------------
class RefRobustCE(nn.CrossEntropyLoss):
    def forward(self, input, target, uncertainty=None):
        if target.ndim == input.ndim:
            assert target.shape[1] == 1
            target = target[:, 0]
        loss = super().forward(input, target.long())
        if uncertainty is not None:
            loss = loss * uncertainty
        return loss.mean()


# ---- utils/seg_utils.py:305-353 (ignore_label=None branch, the only one _build_loss uses) ------------------------
class RefDCAndWeightedCE(nn.Module):
    def __init__(self, soft_dice_kwargs, ce_kwargs, weight_ce=1, weight_dice=1):
        super().__init__()
        self.weight_dice, self.weight_ce = weight_dice, weight_ce
        self.ce = RefRobustCE(**ce_kwargs)
        soft_dice_kwargs = dict(soft_dice_kwargs)
        self.dc = tp.MemoryEfficientSoftDiceLoss(apply_nonlin=tp.softmax_helper_dim1, **soft_dice_kwargs)

    def forward(self, net_output, target, uncertainty=None):
        dc_loss = self.dc(net_output, target, loss_mask=None) if self.weight_dice != 0 else 0
        ce_loss = self.ce(net_output, target[:, 0], uncertainty) if self.weight_ce != 0 else 0
        return self.weight_ce * ce_loss + self.weight_dice * dc_loss


# ---- utils/seg_utils.py:355-372 --------------------------------------------------------------------------------
def ref_build_loss(enable_deep_supervision=False, weight_dice=1):
    loss = RefDCAndWeightedCE({"batch_dice": False, "smooth": 1e-5, "do_bg": False, "ddp": False}, {"reduction": "none"},
                              weight_ce=1, weight_dice=weight_dice)
    if enable_deep_supervision:
        scales = 6
        weights = np.array([1 / (2 ** i) for i in range(scales)])
        weights[-1] = 0
        weights = weights / weights.sum()
        loss = tp.DeepSupervisionWrapper(loss, weights)
    return loss


# ---- train_all.py:519-558 --------------------------------------------------------------------------------------
def ref_joint_step(model_seg, batch, model_sr=None, distiller=None, enable_uncertainty=True):
    """Forward + backward of one stage-2 iteration on CPU (no optimiser step: callers compare losses and gradients).
    `batch` = (img, label_lr, label, uncertainty_lr).  NB `img` is z-scored in place by the teacher sweep, as in the
    reference, so the student sees the normalised image."""
    img, label_lr, label, uncertainty_lr = batch
    loss_lr_obj = ref_build_loss(False, weight_dice=0 if enable_uncertainty else 1)
    loss_hr_obj = ref_build_loss(False, weight_dice=1)
    model_seg.train()
    distill = model_sr is not None and distiller is not None
    if distill:
        with torch.no_grad():
            features_sr = ref_flavr.intermediate_features(model_sr, img, label_lr, normalize=ref_vol.zscore_normalization)
        seg_lr, seg_sr, features_seg = model_seg(img, return_inetermediate_feature=True)
    else:
        seg_lr, seg_sr = model_seg(img)
    if enable_uncertainty:
        loss_lr = loss_lr_obj(seg_lr, label_lr, uncertainty_lr)
        loss_hr = loss_hr_obj(seg_sr, label, None)
    else:
        loss_lr = loss_lr_obj(seg_lr, label_lr)
        loss_hr = loss_hr_obj(seg_sr, label)
    loss = loss_lr + loss_hr
    out = {"loss_lr_seg": loss_lr.detach(), "loss_hr_seg": loss_hr.detach()}
    if distill:
        d = distiller(features_seg[1], features_sr[1])
        loss = loss + d
        out["distill_loss"] = d.detach()
    for p in list(model_seg.parameters()) + (list(distiller.parameters()) if distill else []):
        p.grad = None
    loss.backward()
    out["loss"] = loss.detach()
    return out


# ---------------------------------------------------------------------------------------------------------------
# SR stage: BCEDiceLoss (utils/seg_utils.py:786-886) and the loop body of train_sr (train_all.py:118-139)
# ---------------------------------------------------------------------------------------------------------------
def _ref_flatten(t):
    c = t.size(1)
    return t.permute((1, 0) + tuple(range(2, t.dim()))).contiguous().view(c, -1)


def ref_per_channel_dice(inp, target, epsilon=1e-6):
    """compute_per_channel_dice, utils/seg_utils.py:835-861 (weight=None)."""
    assert inp.size() == target.size()
    inp, target = _ref_flatten(inp), _ref_flatten(target).float()
    intersect = (inp * target).sum(-1)
    denominator = (inp * inp).sum(-1) + (target * target).sum(-1)
    return 2 * (intersect / denominator.clamp(min=epsilon))


class RefBCEDiceLoss(nn.Module):
    def __init__(self, alpha, beta):
        super().__init__()
        self.alpha, self.beta = alpha, beta
        self.bce = nn.BCEWithLogitsLoss()

    def forward(self, inp, target):
        dice = 1. - torch.mean(ref_per_channel_dice(torch.sigmoid(inp), target))
        return self.alpha * self.bce(inp, target) + self.beta * dice


def ref_sr_step(model, patches_lr, patches_hr, loss_obj, loss_seg, slice_separation, num_slices, enable_uncertainty):
    """train_all.py:118-139 up to and including loss.backward(); returns the loss."""
    if num_slices > 1:
        patches_hr = patches_hr[:, :, int(slice_separation) * (num_slices // 2 - 1):int(slice_separation) * (num_slices // 2), ...]
    if enable_uncertainty:
        hat, uncertainty = model(patches_lr)
        loss = loss_obj(hat[:, 0:1, ...], patches_hr[:, 0:1, ...])
        loss += torch.mean(torch.div(torch.abs(hat[:, 0:1, ...] - patches_hr[:, 0:1, ...]), uncertainty) + torch.log(uncertainty))
        error_map = torch.abs(hat[:, 0:1, ...].detach() - patches_hr[:, 0:1, ...])
        loss += loss_obj(uncertainty, error_map)
    else:
        hat = model(patches_lr)
        loss = loss_obj(hat[:, 0:1, ...], patches_hr[:, 0:1, ...])
    loss += loss_seg(hat[:, 1:, ...], patches_hr[:, 1:, ...]) * 1.0
    for p in model.parameters():
        p.grad = None
    loss.backward()
    return loss.detach()
