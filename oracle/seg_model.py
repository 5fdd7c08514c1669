"""CPU restatement (plain PyTorch fp32) of the reference segmentation network -- TEST INFRASTRUCTURE.

Follows models/seg_model.py of the reference:
  * RefDecoder.forward     <- MyUnetDecoder.forward, models/seg_model.py:26-58 (transposed conv, concat with the skip
                              -- upsampled tensor first --, conv stack; segmentation conv on the last stage only unless
                              deep supervision; additionally hands back the last stage's feature map)
  * RefSegModel.__init__   <- SegModel.__init__, models/seg_model.py:154-199 (PlainConvUNet + decoder swap + sr_head:
                              Conv3d(32,16,k3,p1) -> ReLU -> Conv3d(16,num_classes,k5,p2))
  * RefSegModel.forward    <- SegModel.forward, models/seg_model.py:201-210 (trilinear x(upscale,1,1) with
                              align_corners=True on the feature map, then sr_head)
The PlainConvUNet pieces come from oracle/third_party.py (parity unpinned, third-party).  Pinned against the live
reference module by tests/test_oracle_vs_reference.py and the committed tests/golden/segmodel_small.npz.
"""
from __future__ import annotations

import torch
import torch.nn.functional as F
from torch import nn

from . import third_party as tp


class RefDecoder(tp.UNetDecoder):
    def __init__(self, encoder, num_classes, n_conv_per_stage, deep_supervision, deep_features, nonlin_first=False):
        super().__init__(encoder, num_classes, n_conv_per_stage, deep_supervision, nonlin_first)
        self.deep_features = deep_features

    def forward(self, skips):
        low = skips[-1]
        segs, feats = [], []
        last = len(self.stages) - 1
        for s, (up, stage) in enumerate(zip(self.transpconvs, self.stages)):
            low = stage(torch.cat((up(low), skips[-(s + 2)]), dim=1))
            if self.deep_features and s == last:
                feats = low
            if self.deep_supervision:
                segs.append(self.seg_layers[s](low))
            elif s == last:
                segs.append(self.seg_layers[-1](low))
        segs.reverse()
        r = segs if self.deep_supervision else segs[0]
        return (r, feats) if self.deep_features else r


class RefSegModel(tp.PlainConvUNet):
    def __init__(self, input_channels, n_stages, features_per_stage, conv_op, kernel_sizes, strides, n_conv_per_stage,
                 num_classes, upscale, n_conv_per_stage_decoder, conv_bias=False, norm_op=None, norm_op_kwargs=None,
                 dropout_op=None, dropout_op_kwargs=None, nonlin=None, nonlin_kwargs=None, deep_supervision=False,
                 nonlin_first=False):
        super().__init__(input_channels, n_stages, features_per_stage, conv_op, kernel_sizes, strides, n_conv_per_stage,
                         num_classes, n_conv_per_stage_decoder, conv_bias, norm_op, norm_op_kwargs, dropout_op,
                         dropout_op_kwargs, nonlin, nonlin_kwargs, deep_supervision, nonlin_first)
        self.decoder = RefDecoder(self.encoder, num_classes, n_conv_per_stage_decoder, deep_supervision, True, nonlin_first)
        self.upscale = upscale
        self.sr_head = nn.Sequential(nn.Conv3d(32, 16, 3, 1, 1), nn.ReLU(), nn.Conv3d(16, num_classes, 5, 1, 2))

    def forward(self, x, return_inetermediate_feature=False):
        skips = self.encoder(x)
        out, feats = self.decoder(skips)
        hr = self.sr_head(F.interpolate(feats, scale_factor=(self.upscale, 1, 1), mode="trilinear", align_corners=True))
        return (out, hr, skips) if return_inetermediate_feature else (out, hr)


def plan_kwargs(name: str = "3d_fullres") -> dict:
    """Architecture kwargs of the synthetic plans used by the BASELINE configs."""
    common = dict(input_channels=1, conv_op=nn.Conv3d, num_classes=2, upscale=4, conv_bias=True, norm_op=nn.InstanceNorm3d,
                  norm_op_kwargs={"eps": 1e-5, "affine": True}, dropout_op=None, dropout_op_kwargs=None,
                  nonlin=nn.LeakyReLU, nonlin_kwargs={"inplace": True}, deep_supervision=False)
    if name == "3d_fullres":      # config 1 / 3: nnU-Net 3d_fullres defaults
        return dict(n_stages=6, features_per_stage=[32, 64, 128, 256, 320, 320], kernel_sizes=[[3, 3, 3]] * 6,
                    strides=[[1, 1, 1]] + [[2, 2, 2]] * 5, n_conv_per_stage=[2] * 6, n_conv_per_stage_decoder=[2] * 5, **common)
    if name == "anisotropic":     # config 4: the plan the Distiller's shapes require
        return dict(n_stages=6, features_per_stage=[32, 64, 128, 256, 320, 320],
                    kernel_sizes=[[1, 3, 3], [1, 3, 3]] + [[3, 3, 3]] * 4,
                    strides=[[1, 1, 1], [1, 2, 2], [1, 2, 2], [2, 2, 2], [2, 2, 2], [1, 2, 2]],
                    n_conv_per_stage=[2] * 6, n_conv_per_stage_decoder=[2] * 5, **common)
    if name == "tiny":            # 3 stages, for second-scale CPU tests
        return dict(n_stages=3, features_per_stage=[32, 64, 128], kernel_sizes=[[3, 3, 3]] * 3,
                    strides=[[1, 1, 1], [2, 2, 2], [2, 2, 2]], n_conv_per_stage=[2] * 3, n_conv_per_stage_decoder=[2] * 2,
                    **common)
    raise KeyError(name)


def build(name: str = "3d_fullres", seed: int = 1234) -> RefSegModel:
    torch.manual_seed(seed)
    return RefSegModel(**plan_kwargs(name))
