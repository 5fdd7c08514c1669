"""Restatements of the un-vendored third-party symbols the reference's hot path executes.  PARITY UNPINNED: the
packages are absent from /root/reference and from this image (no network); nothing in the reference pins them.
Each function follows the published behaviour of the pinned version (requirements.txt) as summarised in
SURVEY.md section 8(c).  TEST INFRASTRUCTURE -- never imported by rehrseg_b200/.

  dynamic_network_architectures==0.3.1 (requirements.txt:20) -> ConvDropoutNormReLU, StackedConvBlocks,
      PlainConvEncoder, UNetDecoder, PlainConvUNet      (imported at models/seg_model.py:9-10)
  nnunetv2==2.3.1 (requirements.txt:53)                -> compute_gaussian, MemoryEfficientSoftDiceLoss,
      SoftDiceLoss, softmax_helper_dim1                 (imported at utils/seg_utils.py:9-14)
  acvl_utils==0.2 (requirements.txt:2)                  -> pad_nd_image (imported at utils/seg_utils.py:7)
"""
from __future__ import annotations

from functools import lru_cache
from typing import List, Sequence, Tuple, Union

import numpy as np
import torch
from scipy.ndimage import gaussian_filter
from torch import nn


def _as_list(v, n):
    return [v] * n if isinstance(v, int) else list(v)


def _dim_of(conv_op) -> int:
    return {nn.Conv1d: 1, nn.Conv2d: 2, nn.Conv3d: 3}[conv_op]


def _transp_of(conv_op):
    return {nn.Conv1d: nn.ConvTranspose1d, nn.Conv2d: nn.ConvTranspose2d, nn.Conv3d: nn.ConvTranspose3d}[conv_op]


class ConvDropoutNormReLU(nn.Module):
    """conv(k, stride, padding=(k-1)//2, dilation 1, bias) -> [dropout] -> [norm] -> [nonlin]; sub-modules are registered
    as .conv/.dropout/.norm/.nonlin AND again inside .all_modules (hence the duplicated state_dict keys)."""

    def __init__(self, conv_op, input_channels, output_channels, kernel_size, stride, conv_bias=False, norm_op=None,
                 norm_op_kwargs=None, dropout_op=None, dropout_op_kwargs=None, nonlin=None, nonlin_kwargs=None,
                 nonlin_first=False):
        super().__init__()
        dim = _dim_of(conv_op)
        kernel_size = _as_list(kernel_size, dim)
        self.stride = _as_list(stride, dim)
        self.input_channels, self.output_channels = input_channels, output_channels
        ops = []
        self.conv = conv_op(input_channels, output_channels, kernel_size, self.stride,
                            padding=[(k - 1) // 2 for k in kernel_size], dilation=1, bias=conv_bias)
        ops.append(self.conv)
        if dropout_op is not None:
            self.dropout = dropout_op(**(dropout_op_kwargs or {}))
            ops.append(self.dropout)
        if norm_op is not None:
            self.norm = norm_op(output_channels, **(norm_op_kwargs or {}))
            ops.append(self.norm)
        if nonlin is not None:
            self.nonlin = nonlin(**(nonlin_kwargs or {}))
            ops.append(self.nonlin)
        if nonlin_first and norm_op is not None and nonlin is not None:
            ops[-1], ops[-2] = ops[-2], ops[-1]
        self.all_modules = nn.Sequential(*ops)

    def forward(self, x):
        return self.all_modules(x)


class StackedConvBlocks(nn.Module):
    def __init__(self, num_convs, conv_op, input_channels, output_channels, kernel_size, initial_stride, conv_bias=False,
                 norm_op=None, norm_op_kwargs=None, dropout_op=None, dropout_op_kwargs=None, nonlin=None,
                 nonlin_kwargs=None, nonlin_first=False):
        super().__init__()
        outs = _as_list(output_channels, num_convs)
        common = (conv_bias, norm_op, norm_op_kwargs, dropout_op, dropout_op_kwargs, nonlin, nonlin_kwargs, nonlin_first)
        blocks = [ConvDropoutNormReLU(conv_op, input_channels, outs[0], kernel_size, initial_stride, *common)]
        blocks += [ConvDropoutNormReLU(conv_op, outs[i - 1], outs[i], kernel_size, 1, *common) for i in range(1, num_convs)]
        self.convs = nn.Sequential(*blocks)
        self.output_channels = outs[-1]
        self.initial_stride = _as_list(initial_stride, _dim_of(conv_op))

    def forward(self, x):
        return self.convs(x)


class PlainConvEncoder(nn.Module):
    def __init__(self, input_channels, n_stages, features_per_stage, conv_op, kernel_sizes, strides, n_conv_per_stage,
                 conv_bias=False, norm_op=None, norm_op_kwargs=None, dropout_op=None, dropout_op_kwargs=None, nonlin=None,
                 nonlin_kwargs=None, return_skips=False, nonlin_first=False, pool="conv"):
        super().__init__()
        assert pool == "conv", "the reference never selects a pooling variant"
        kernel_sizes = [kernel_sizes] * n_stages if isinstance(kernel_sizes, int) else list(kernel_sizes)
        features_per_stage = _as_list(features_per_stage, n_stages)
        n_conv_per_stage = _as_list(n_conv_per_stage, n_stages)
        strides = [strides] * n_stages if isinstance(strides, int) else list(strides)
        stages = []
        cin = input_channels
        for s in range(n_stages):
            stages.append(nn.Sequential(StackedConvBlocks(
                n_conv_per_stage[s], conv_op, cin, features_per_stage[s], kernel_sizes[s], strides[s], conv_bias, norm_op,
                norm_op_kwargs, dropout_op, dropout_op_kwargs, nonlin, nonlin_kwargs, nonlin_first)))
            cin = features_per_stage[s]
        self.stages = nn.Sequential(*stages)
        self.output_channels = features_per_stage
        self.strides = [_as_list(i, _dim_of(conv_op)) for i in strides]
        self.return_skips = return_skips
        self.conv_op, self.norm_op, self.norm_op_kwargs = conv_op, norm_op, norm_op_kwargs
        self.nonlin, self.nonlin_kwargs = nonlin, nonlin_kwargs
        self.dropout_op, self.dropout_op_kwargs = dropout_op, dropout_op_kwargs
        self.conv_bias, self.kernel_sizes = conv_bias, kernel_sizes

    def forward(self, x):
        ret = []
        for s in self.stages:
            x = s(x)
            ret.append(x)
        return ret if self.return_skips else ret[-1]


class UNetDecoder(nn.Module):
    def __init__(self, encoder, num_classes, n_conv_per_stage, deep_supervision, nonlin_first=False):
        super().__init__()
        self.deep_supervision = deep_supervision
        self.encoder = encoder
        self.num_classes = num_classes
        n_enc = len(encoder.output_channels)
        n_conv_per_stage = _as_list(n_conv_per_stage, n_enc - 1)
        transp = _transp_of(encoder.conv_op)
        stages, transpconvs, seg_layers = [], [], []
        for s in range(1, n_enc):
            below, skip = encoder.output_channels[-s], encoder.output_channels[-(s + 1)]
            stride = encoder.strides[-s]
            transpconvs.append(transp(below, skip, stride, stride, bias=encoder.conv_bias))
            stages.append(StackedConvBlocks(n_conv_per_stage[s - 1], encoder.conv_op, 2 * skip, skip,
                                            encoder.kernel_sizes[-(s + 1)], 1, encoder.conv_bias, encoder.norm_op,
                                            encoder.norm_op_kwargs, encoder.dropout_op, encoder.dropout_op_kwargs,
                                            encoder.nonlin, encoder.nonlin_kwargs, nonlin_first))
            seg_layers.append(encoder.conv_op(skip, num_classes, 1, 1, 0, bias=True))
        self.stages = nn.ModuleList(stages)
        self.transpconvs = nn.ModuleList(transpconvs)
        self.seg_layers = nn.ModuleList(seg_layers)

    def forward(self, skips):
        lres = skips[-1]
        outs = []
        for s in range(len(self.stages)):
            x = self.stages[s](torch.cat((self.transpconvs[s](lres), skips[-(s + 2)]), 1))
            if self.deep_supervision:
                outs.append(self.seg_layers[s](x))
            elif s == len(self.stages) - 1:
                outs.append(self.seg_layers[-1](x))
            lres = x
        outs = outs[::-1]
        return outs if self.deep_supervision else outs[0]


class PlainConvUNet(nn.Module):
    def __init__(self, input_channels, n_stages, features_per_stage, conv_op, kernel_sizes, strides, n_conv_per_stage,
                 num_classes, n_conv_per_stage_decoder, conv_bias=False, norm_op=None, norm_op_kwargs=None, dropout_op=None,
                 dropout_op_kwargs=None, nonlin=None, nonlin_kwargs=None, deep_supervision=False, nonlin_first=False):
        super().__init__()
        self.encoder = PlainConvEncoder(input_channels, n_stages, features_per_stage, conv_op, kernel_sizes, strides,
                                        _as_list(n_conv_per_stage, n_stages), conv_bias, norm_op, norm_op_kwargs, dropout_op,
                                        dropout_op_kwargs, nonlin, nonlin_kwargs, return_skips=True, nonlin_first=nonlin_first)
        self.decoder = UNetDecoder(self.encoder, num_classes, _as_list(n_conv_per_stage_decoder, n_stages - 1),
                                   deep_supervision, nonlin_first=nonlin_first)

    def forward(self, x):
        return self.decoder(self.encoder(x))


# ---------------------------------------------------------------------------------------------------------------
# nnunetv2 2.3.1: sliding-window Gaussian importance map
# ---------------------------------------------------------------------------------------------------------------
@lru_cache(maxsize=2)
def compute_gaussian(tile_size: Tuple[int, ...], sigma_scale: float = 1. / 8, value_scaling_factor: float = 1,
                     dtype=torch.float16, device=torch.device("cpu")) -> torch.Tensor:
    """Delta at the tile centre -> scipy gaussian_filter(sigma = tile*sigma_scale, zero-padded) -> scaled so the
    maximum equals value_scaling_factor -> cast to `dtype` -> zeros replaced by the smallest non-zero entry."""
    tmp = np.zeros(tile_size)
    tmp[tuple(i // 2 for i in tile_size)] = 1
    g = gaussian_filter(tmp, [i * sigma_scale for i in tile_size], 0, mode="constant", cval=0)
    g = torch.from_numpy(g)
    g = g / torch.max(g) * value_scaling_factor
    g = g.to(device=device, dtype=dtype)
    g[g == 0] = torch.min(g[g != 0])
    return g


# ---------------------------------------------------------------------------------------------------------------
# acvl_utils 0.2: pad_nd_image
# ---------------------------------------------------------------------------------------------------------------
def pad_nd_image(image, new_shape=None, mode="constant", kwargs=None, return_slicer=False, shape_must_be_divisible_by=None):
    """Pad the trailing len(new_shape) dims of `image` up to new_shape (never crops): below = diff // 2,
    above = diff // 2 + diff % 2.  Returns (padded, slicer) when return_slicer."""
    kwargs = kwargs or {}
    old = np.array(image.shape)
    if new_shape is None:
        new_shape = old
    nd = len(new_shape)
    tgt = old.copy()
    tgt[-nd:] = np.maximum(old[-nd:], np.asarray(new_shape))
    if shape_must_be_divisible_by is not None:
        div = _as_list(shape_must_be_divisible_by, nd)
        for i in range(nd):
            ax = len(old) - nd + i
            if tgt[ax] % div[i] != 0:
                tgt[ax] += div[i] - tgt[ax] % div[i]
    diff = tgt - old
    below = diff // 2
    above = diff // 2 + diff % 2
    pads = [[int(b), int(a)] for b, a in zip(below, above)]
    if any(b or a for b, a in pads):
        if isinstance(image, torch.Tensor):
            flat = [v for p in pads[::-1] for v in p]
            res = torch.nn.functional.pad(image, flat, mode=mode, **kwargs)
        else:
            np_kwargs = {"constant_values": kwargs["value"]} if (mode == "constant" and "value" in kwargs) else {}
            res = np.pad(image, pads, mode, **np_kwargs)
    else:
        res = image
    if not return_slicer:
        return res
    slicer = tuple(slice(p[0], res.shape[i] - p[1]) for i, p in enumerate(pads))
    return res, slicer


# ---------------------------------------------------------------------------------------------------------------
# nnunetv2 2.3.1: dice losses / helpers (needed only so the reference's _build_loss runs for config 4)
# ---------------------------------------------------------------------------------------------------------------
def softmax_helper_dim1(x: torch.Tensor) -> torch.Tensor:
    return torch.softmax(x, 1)


class MemoryEfficientSoftDiceLoss(nn.Module):
    def __init__(self, apply_nonlin=None, batch_dice=False, do_bg=True, smooth=1., ddp=True):
        super().__init__()
        self.apply_nonlin, self.batch_dice, self.do_bg, self.smooth, self.ddp = apply_nonlin, batch_dice, do_bg, smooth, ddp

    def forward(self, x, y, loss_mask=None):
        if self.apply_nonlin is not None:
            x = self.apply_nonlin(x)
        axes = tuple(range(2, x.ndim))
        with torch.no_grad():
            if x.ndim != y.ndim:
                y = y.view((y.shape[0], 1, *y.shape[1:]))
            if x.shape == y.shape:
                y_onehot = y
            else:
                y_onehot = torch.zeros(x.shape, device=x.device, dtype=torch.bool)
                y_onehot.scatter_(1, y.long(), 1)
            if not self.do_bg:
                y_onehot = y_onehot[:, 1:]
            sum_gt = y_onehot.sum(axes) if loss_mask is None else (y_onehot * loss_mask).sum(axes)
        if not self.do_bg:
            x = x[:, 1:]
        if loss_mask is None:
            intersect = (x * y_onehot).sum(axes)
            sum_pred = x.sum(axes)
        else:
            intersect = (x * y_onehot * loss_mask).sum(axes)
            sum_pred = (x * loss_mask).sum(axes)
        if self.batch_dice:
            intersect, sum_pred, sum_gt = intersect.sum(0), sum_pred.sum(0), sum_gt.sum(0)
        dc = (2 * intersect + self.smooth) / (torch.clip(sum_gt + sum_pred + self.smooth, 1e-8))
        return -dc.mean()


class SoftDiceLoss(MemoryEfficientSoftDiceLoss):
    """Only used as a default-argument value by the reference (utils/seg_utils.py:306); same formula."""


class DeepSupervisionWrapper(nn.Module):
    def __init__(self, loss, weight_factors=None):
        super().__init__()
        self.loss, self.weight_factors = loss, weight_factors

    def forward(self, *args):
        w = self.weight_factors if self.weight_factors is not None else (1,) * len(args[0])
        return sum(w[i] * self.loss(*inputs) for i, inputs in enumerate(zip(*args)) if w[i] != 0.0)
