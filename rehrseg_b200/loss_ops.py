"""Fused CUDA reductions for the stage-2 losses (csrc/loss_ops.cu) as autograd Functions, and the loss modules built on them.

Opt-in counterparts of the plain-PyTorch mirrors in `train_step.py` (same constructors, same values):
  * `FusedDCAndWeightedCELoss` <- `DC_and_weighted_CE_loss` as `_build_loss` configures it (utils/seg_utils.py:305-372)
  * `FusedDistiller`           <- `Distiller` (models/seg_model.py:115-151)
Each kernel makes one pass over the fp32 NCDHW tensor and returns per-(sample, class/channel) SUMS; the handful of scalar
operations that turn sums into the loss stay ordinary autograd ops on tiny tensors, and the backward kernels consume the
gradients of the sums.  CUDA only: like every engine op they raise `RehrError` on a CPU tensor or a missing library.
"""
from __future__ import annotations

import ctypes as C
from typing import Optional

import torch
import torch.nn.functional as F
from torch import nn

from . import _lib as L
from ._lib import check, lib, ptr, stream_ptr


def _f32c(t: torch.Tensor, what: str) -> torch.Tensor:
    if not t.is_cuda:
        raise L.RehrError(f"{what}: expected a CUDA tensor (rehrseg_b200 has no CPU path)")
    t = t if t.dtype == torch.float32 else t.float()
    return t if t.is_contiguous() else t.contiguous()


class SegLossSums(torch.autograd.Function):
    """(logits [B,C,*], target [B,1,*] class indices, weight [*] or None) ->
    ce_sum [B] = sum_v weight*CE, inter [B,C] = sum p*[y=c], pred [B,C] = sum p, count [B,C] = sum [y=c]   (p = softmax)."""

    @staticmethod
    def forward(ctx, logits, target, weight):
        lg = _f32c(logits.detach(), "seg_loss_sums")
        tg = _f32c(target.detach(), "seg_loss_sums")
        wt = _f32c(weight.detach(), "seg_loss_sums") if weight is not None else None
        b, c = lg.shape[0], lg.shape[1]
        v = lg[0, 0].numel()
        if tg.numel() != b * v or (wt is not None and wt.numel() != v):
            raise L.RehrError("seg_loss_sums: target must be [B,1,*] and weight [*] matching the logits")
        blocks = lib().rehr_loss_blocks(v)
        k = 1 + 3 * c
        partial = torch.empty((b, blocks, k), dtype=torch.float32, device=lg.device)
        check(lib().rehr_seg_loss_sums(ptr(lg), ptr(tg), ptr(wt), b, c, v, ptr(partial), stream_ptr()), "seg_loss_sums")
        sums = partial.sum(1)
        ctx.save_for_backward(lg, tg, wt)
        per = sums[:, 1:].reshape(b, c, 3)
        count = per[:, :, 2]
        ctx.mark_non_differentiable(count)
        return sums[:, 0], per[:, :, 0], per[:, :, 1], count

    @staticmethod
    def backward(ctx, g_ce, g_int, g_pred, _g_cnt):
        lg, tg, wt = ctx.saved_tensors
        b, c = lg.shape[0], lg.shape[1]
        v = lg[0, 0].numel()
        dl = torch.empty_like(lg)
        z = lambda g, shape: (g if g is not None else torch.zeros(shape, device=lg.device)).float().contiguous()  # noqa: E731
        check(lib().rehr_seg_loss_bwd(ptr(lg), ptr(tg), ptr(wt), b, c, v, ptr(z(g_ce, (b,))), ptr(z(g_int, (b, c))),
                                      ptr(z(g_pred, (b, c))), ptr(dl), stream_ptr()), "seg_loss_bwd")
        return dl, None, None


class CosineSums(torch.autograd.Function):
    """(a [B,C,*], b [B,C,*]) -> (sum_v ahat*bhat, sum_v ahat^2, sum_v bhat^2), each [B,C]; xhat = F.normalize(x, dim=1).
    Gradient flows to `a` only (the teacher map is detached in the reference)."""

    @staticmethod
    def forward(ctx, a, b):
        a32, b32 = _f32c(a.detach(), "cosine_sums"), _f32c(b.detach(), "cosine_sums")
        if a32.shape != b32.shape:
            raise L.RehrError("cosine_sums: shape mismatch")
        n, c = a32.shape[0], a32.shape[1]
        v = a32[0, 0].numel()
        blocks = lib().rehr_loss_blocks(v)
        partial = torch.empty((n, blocks, 3, c), dtype=torch.float32, device=a32.device)
        check(lib().rehr_cosine_sums(ptr(a32), ptr(b32), n, c, v, ptr(partial), stream_ptr()), "cosine_sums")
        s = partial.sum(1)
        ctx.save_for_backward(a32, b32)
        ctx.adtype = a.dtype
        return s[:, 0], s[:, 1], s[:, 2]

    @staticmethod
    def backward(ctx, g_ab, g_aa, _g_bb):
        a32, b32 = ctx.saved_tensors
        n, c = a32.shape[0], a32.shape[1]
        v = a32[0, 0].numel()
        z = lambda g: (g if g is not None else torch.zeros((n, c), device=a32.device)).float().contiguous()  # noqa: E731
        da = torch.empty_like(a32)
        check(lib().rehr_cosine_sums_bwd(ptr(a32), ptr(b32), n, c, v, ptr(z(g_ab)), ptr(z(g_aa)), ptr(da), stream_ptr()),
              "cosine_sums_bwd")
        return da.to(ctx.adtype), None


class PlaneMaxPool(torch.autograd.Function):
    """x [N,C,S,H,W] -> [(N S), C, OH, OW]: per slice, max over non-overlapping ph x pw windows, ceil mode -- the
    `rearrange 'b c s h w -> (b s) c h w'` + `nn.MaxPool2d(k, stride=k, ceil_mode=True)` of models/seg_model.py:103-110."""

    @staticmethod
    def forward(ctx, x, ph, pw):
        x32 = _f32c(x.detach(), "plane_maxpool")
        n, c, s, h, w = x32.shape
        oh, ow = -(-h // ph), -(-w // pw)
        out = torch.empty((n * s, c, oh, ow), dtype=torch.float32, device=x32.device)
        idx = torch.empty((n * s, c, oh, ow), dtype=torch.int32, device=x32.device)
        check(lib().rehr_plane_maxpool(ptr(x32), n, c, s, h, w, int(ph), int(pw), ptr(out), ptr(idx), stream_ptr()), "plane_maxpool")
        ctx.save_for_backward(idx)
        ctx.cfg = (n, c, s, h, w, int(ph), int(pw), x.dtype)
        return out

    @staticmethod
    def backward(ctx, g):
        (idx,) = ctx.saved_tensors
        n, c, s, h, w, ph, pw, dt = ctx.cfg
        dx = torch.zeros((n, c, s, h, w), dtype=torch.float32, device=g.device)
        check(lib().rehr_plane_maxpool_bwd(ptr(g.float().contiguous()), ptr(idx), n, c, s, h, w, ph, pw, ptr(dx), stream_ptr()),
              "plane_maxpool_bwd")
        return dx.to(dt), None, None


# ---------------------------------------------------------------------------------------------------------------
# loss modules
# ---------------------------------------------------------------------------------------------------------------
class FusedDCAndWeightedCELoss(nn.Module):
    """`train_step.DCAndWeightedCELoss` on the fused sums.  The reference multiplies the CE map [B,D,H,W] by the uncertainty
    [B,1,D,H,W] (utils/seg_utils.py:299-301,349), which broadcasts over the batch: mean over [B,B,V] of ce[b']*unc[b] equals
    the mean over [B,V] of ce[b'] * mean_b unc[b] -- the per-voxel weight handed to the kernel."""

    def __init__(self, weight_ce: float = 1, weight_dice: float = 1, smooth: float = 1e-5):
        super().__init__()
        self.weight_ce, self.weight_dice, self.smooth = weight_ce, weight_dice, smooth

    def forward(self, net_output: torch.Tensor, target: torch.Tensor, uncertainty: Optional[torch.Tensor] = None):
        assert target.shape[1] == 1, "target must be [B,1,...] label indices"
        b = net_output.shape[0]
        v = net_output[0, 0].numel()
        w = uncertainty.float().mean(0)[0] if (uncertainty is not None and self.weight_ce != 0) else None
        ce_sum, inter, pred, count = SegLossSums.apply(net_output, target, w)
        total = 0
        if self.weight_dice != 0:
            dc = (2 * inter[:, 1:] + self.smooth) / torch.clip(count[:, 1:] + pred[:, 1:] + self.smooth, 1e-8)
            total = total + self.weight_dice * (-dc.mean())
        if self.weight_ce != 0:
            total = total + self.weight_ce * ce_sum.sum() / (b * v)
        return total


def fused_cosine_distance_loss(a: torch.Tensor, b: torch.Tensor) -> torch.Tensor:
    """models/seg_model.py:60-78 on the fused sums."""
    ab, aa, bb = CosineSums.apply(a, b)
    cos = ab / (aa.sqrt().clamp_min(1e-8) * bb.sqrt().clamp_min(1e-8))
    return (1 - cos).mean()


def fused_structure_loss(student: torch.Tensor, teacher: torch.Tensor, scale: float = 0.5) -> torch.Tensor:
    """models/seg_model.py:80-113 with the two max-pools on the fused kernel; the 4x4 Gram matrices stay PyTorch."""
    from .train_step import _gram
    _, _, s, th, tw = student.shape
    ph, pw = int(th * scale), int(tw * scale)
    ps = PlaneMaxPool.apply(student, ph, pw)
    pt = PlaneMaxPool.apply(teacher.detach(), ph, pw)
    err = (_gram(pt) - _gram(ps)).pow(2) / float((pt.shape[-1] * pt.shape[-2]) ** 2) / pt.shape[0]
    return err.sum() / s


class FusedDistiller(nn.Module):
    """`train_step.Distiller` (same constructor / state_dict) with the cosine and structure terms on the fused kernels."""

    def __init__(self, student_dim, teacher_dim, lambda_l1=0.0, lambda_cosine=0.0, lambda_structure=0.0):
        super().__init__()
        self.lambda_l1, self.lambda_cosine, self.lambda_structure = lambda_l1, lambda_cosine, lambda_structure
        self.distill = nn.Conv3d(student_dim, teacher_dim, kernel_size=1, stride=1, padding=0)

    def forward(self, feature_student, feature_teacher):
        loss = 0
        if self.lambda_structure > 0:
            loss = loss + self.lambda_structure * fused_structure_loss(feature_student, feature_teacher, 0.5)
        projected = self.distill(feature_student)
        if self.lambda_l1 > 0:
            loss = loss + self.lambda_l1 * F.smooth_l1_loss(projected, feature_teacher)
        if self.lambda_cosine > 0:
            loss = loss + self.lambda_cosine * fused_cosine_distance_loss(projected, feature_teacher)
        return loss


def build_fused_loss(enable_deep_supervision: bool = False, weight_dice: float = 1) -> nn.Module:
    """`_build_loss` (utils/seg_utils.py:355-372) on the fused kernels; deep supervision is off (train_all.py:471)."""
    if enable_deep_supervision:
        raise NotImplementedError("train_all.py:471 fixes enable_deep_supervision=False")
    return FusedDCAndWeightedCELoss(weight_ce=1, weight_dice=weight_dice, smooth=1e-5)
