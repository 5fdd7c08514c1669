"""CPU restatement (plain PyTorch fp32) of the reference FLAVR self-SR network -- TEST INFRASTRUCTURE.

  RefEncoder      <- VideoResNet / BasicBlock / SEGating / Conv3DSimple / BasicStem as instantiated by unet_18 with
                     batchnorm=identity and useBias=True (models/FLAVR/resnet_3D.py:19-50,100-261)
  RefFLAVR        <- UNet_3D_3D (models/FLAVR/FLAVR_arch.py:117-248): decoder of Conv_3d / upConv3D + SE gates with
                     LeakyReLU(0.2) and skip concatenation, depth unbound into channels, 2-D fuse conv, then either the
                     reflect-padded 7x7 out-conv + tanh(img + mean) (plain head) or the 16-expert softmax mixture + sigmoid
                     uncertainty (UASR head).  forward mutates images[:, 0:1] in place exactly like the reference.
  intermediate_features <- get_intermediate_features (train_all.py:85-112), device-agnostic
Parameter names equal the reference's, so `load_state_dict` of a reference / product state works.  Pinned against the live
reference module by tests/test_oracle_vs_reference.py and tests/golden/flavr_small.npz (oracle/make_golden.py).
"""
from __future__ import annotations

import torch
import torch.nn.functional as F
from torch import nn


class _Id(nn.Module):
    def forward(self, x):
        return x


class _SE(nn.Module):
    def __init__(self, c):
        super().__init__()
        self.pool = nn.AdaptiveAvgPool3d(1)
        self.attn_layer = nn.Sequential(nn.Conv3d(c, c, 1, bias=True), nn.Sigmoid())

    def forward(self, x):
        return x * self.attn_layer(self.pool(x))


class _Block(nn.Module):
    def __init__(self, cin, cout, stride=1, downsample=None):
        super().__init__()
        self.conv1 = nn.Sequential(nn.Conv3d(cin, cout, 3, stride, 1, bias=True), _Id(), nn.ReLU(inplace=True))
        self.conv2 = nn.Sequential(nn.Conv3d(cout, cout, 3, 1, 1, bias=True), _Id())
        self.fg = _SE(cout)
        self.downsample = downsample

    def forward(self, x):
        y = self.fg(self.conv2(self.conv1(x)))
        return F.relu(y + (x if self.downsample is None else self.downsample(x)))


class RefEncoder(nn.Module):
    def __init__(self, img_channels):
        super().__init__()
        self.stem = nn.Sequential(nn.Conv3d(img_channels, 64, (3, 7, 7), (1, 2, 2), (1, 3, 3), bias=True), _Id(), nn.ReLU())
        cfg = [(64, 64, (1, 1, 1)), (64, 128, (1, 2, 2)), (128, 256, (1, 2, 2)), (256, 512, (1, 1, 1))]
        for i, (cin, cout, st) in enumerate(cfg, 1):
            ds = None
            if cin != cout or st != (1, 1, 1):
                ds = nn.Sequential(nn.Conv3d(cin, cout, 1, st, bias=False), _Id())
            setattr(self, f"layer{i}", nn.Sequential(_Block(cin, cout, st if ds is not None else 1, ds), _Block(cout, cout)))
        for m in self.modules():
            if isinstance(m, nn.Conv3d):
                nn.init.kaiming_normal_(m.weight, mode="fan_out", nonlinearity="relu")
                if m.bias is not None:
                    nn.init.zeros_(m.bias)

    def forward(self, x):
        x0 = self.stem(x)
        x1 = self.layer1(x0)
        x2 = self.layer2(x1)
        x3 = self.layer3(x2)
        return x0, x1, x2, x3, self.layer4(x3)


class _C3(nn.Module):
    def __init__(self, cin, cout):
        super().__init__()
        self.conv = nn.Sequential(nn.Conv3d(cin, cout, 3, 1, 1, bias=True), _SE(cout))

    def forward(self, x):
        return self.conv(x)


class _Up(nn.Module):
    def __init__(self, cin, cout):
        super().__init__()
        self.upconv = nn.Sequential(nn.ConvTranspose3d(cin, cout, (3, 4, 4), (1, 2, 2), (1, 1, 1)), _SE(cout))

    def forward(self, x):
        return self.upconv(x)


class _C2(nn.Module):
    def __init__(self, cin, cout, k, pad=0):
        super().__init__()
        self.conv = nn.Sequential(nn.Conv2d(cin, cout, k, 1, pad, bias=True))

    def forward(self, x):
        return self.conv(x)


class RefFLAVR(nn.Module):
    def __init__(self, img_channels=2, n_inputs=4, n_outputs=4, use_uncertainty=False):
        super().__init__()
        self.img_channels, self.n_inputs, self.n_outputs, self.use_uncertainty = img_channels, n_inputs, n_outputs, use_uncertainty
        self.encoder = RefEncoder(img_channels)
        self.decoder = nn.Sequential(_C3(512, 256), _Up(512, 128), _Up(256, 64), _C3(128, 64), _Up(128, 64))
        self.feature_fuse = _C2(64 * n_inputs, 64 * n_inputs if use_uncertainty else 64, 3, 1)
        self.feature_fuse1 = _C2(64 * n_inputs, 64 * img_channels, 1)
        if use_uncertainty:
            self.uncertainty_early = _C2(64 * n_inputs, 64, 1)
            self.uncertainty_out = nn.Conv3d(64 // n_outputs, 1, 1)
        self.outconv = nn.Sequential(nn.ReflectionPad2d(3), nn.Conv2d(64, img_channels * n_outputs, 7))

    def forward(self, images, return_inetermediate_uncertainty=False, return_inetermediate_feature=False):
        mean_ = images[:, 0:1].mean((2, 3, 4), keepdim=True)
        images[:, 0:1] = images[:, 0:1] - mean_
        x0, x1, x2, x3, x4 = self.encoder(images)
        if return_inetermediate_feature:
            return x0, x1, x2, x3, x4
        act = lambda t: F.leaky_relu(t, 0.2)
        d = torch.cat([act(self.decoder[0](x4)), x3], 1)
        d = torch.cat([act(self.decoder[1](d)), x2], 1)
        d = torch.cat([act(self.decoder[2](d)), x1], 1)
        d = torch.cat([act(self.decoder[3](d)), x0], 1)
        d = act(self.decoder[4](d))
        d = torch.cat(torch.unbind(d, 2), 1)
        if self.use_uncertainty:
            d = act(self.feature_fuse(d))
            out = torch.stack(torch.chunk(self.feature_fuse1(d), self.n_outputs, 1), 2)
            sm = torch.softmax(torch.stack(torch.chunk(self.uncertainty_early(d), self.n_outputs, 1), 2), 1)
            if return_inetermediate_uncertainty:
                n = sm.shape[1]
                return ([(torch.tanh(out[:, 2 * i:2 * i + 1]) + 1) / 2 for i in range(n)], [sm[:, i:i + 1] for i in range(n)],
                        [out[:, 2 * i + 1:2 * i + 2] for i in range(n)])
            res = 0
            for i in range(sm.shape[1]):
                res = res + torch.cat([(torch.tanh(out[:, 2 * i:2 * i + 1]) + 1) / 2 * sm[:, i:i + 1],
                                       out[:, 2 * i + 1:2 * i + 2] * sm[:, i:i + 1]], 1)
            return res, torch.sigmoid(self.uncertainty_out(sm))
        o = self.outconv(act(self.feature_fuse(d)))
        m2 = mean_.squeeze(2)
        parts = torch.split(o, self.img_channels, 1)
        if self.img_channels > 1:
            parts = [torch.cat([torch.tanh(p[:, 0:1] + m2), p[:, 1:2]], 1) for p in parts]
        else:
            parts = [p + m2 for p in parts]
        return torch.stack(parts, 2)


def build(use_uncertainty=False, seed=1234, img_channels=2) -> RefFLAVR:
    torch.manual_seed(seed)
    return RefFLAVR(img_channels, 4, 4, use_uncertainty)


def intermediate_features(model, img_lr, label_lr, normalize=None):
    """train_all.py:85-112.  One encoder pass per 4-slice window (zero-padded at both ends); slice 1 of every window and
    slice 2 of the last one are stitched along D."""
    if normalize is not None:
        img_lr = normalize(img_lr)
    x = torch.cat((img_lr, label_lr), 1)
    depth = x.shape[2]
    keep = {}
    feats = None
    for st in range(depth - 1):
        if st == 0:
            w = x[:, :, 0:3]
            w = torch.cat([w.new_zeros(w.shape[0], w.shape[1], 4 - w.shape[2], *w.shape[3:]), w], 2)
        elif st == depth - 2:
            w = x[:, :, st - 1:]
            w = torch.cat([w, w.new_zeros(w.shape[0], w.shape[1], 4 - w.shape[2], *w.shape[3:])], 2)
        else:
            w = x[:, :, st - 1:st + 3]
        feats = model(w.clone(), return_inetermediate_feature=True)
        for i, f in enumerate(feats):
            keep.setdefault(i, []).append(f[:, :, 1:2])
    for i, f in enumerate(feats):
        keep[i].append(f[:, :, 2:3])
    return {i: torch.cat(v, 2) for i, v in keep.items()}
