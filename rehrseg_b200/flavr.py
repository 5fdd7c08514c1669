"""B200-native FLAVR `UNet_3D_3D` self-SR network -- drop-in for models/FLAVR/FLAVR_arch.py:117-248 and
models/FLAVR/resnet_3D.py (R3D-18 encoder with SE "feature gating", no BatchNorm) of the reference.

* ``UNet_3D_3D(img_channels, block, n_inputs, n_outputs, batchnorm, joinType, upmode, use_uncertainty)`` -- same constructor
  signature (FLAVR_arch.py:118; called at train_all.py:336-345,401-410), same module tree and construction / init order,
  hence the same ``state_dict`` keys (``encoder.stem.0``, ``encoder.layerL.B.conv{1,2}.0``, ``.fg.attn_layer.0``,
  ``.downsample.0``, ``decoder.N.{conv,upconv}.{0,1.attn_layer.0}``, ``feature_fuse.conv.0``, ``feature_fuse1.conv.0``,
  ``uncertainty_early.conv.0``, ``uncertainty_out``, ``outconv.1``) and the same default-init weights for the same seed.
* ``convert(model)`` -- re-route an already built reference ``UNet_3D_3D`` through the engine, keeping its Parameters.
* ``forward(images, return_inetermediate_uncertainty=False, return_inetermediate_feature=False)`` -- the reference's
  signature, return structure and its in-place mean subtraction on the caller's tensor (FLAVR_arch.py:180-181).
* ``apply_to_vol_flavr`` / ``get_intermediate_features`` -- the window sweeps of utils/sr_utils.py:102-135 and
  train_all.py:85-112 with the windows BATCHED through the network instead of one forward per window.

Only the configuration the reference uses is implemented (block="unet_18", batchnorm=False, joinType="concat",
upmode="transpose"); anything else raises RehrError.  Convolutions, SE gates and the overlapping transposed convolutions
run on the sm_100a engine in channels-last bf16; the 16-expert UASR mixture (a few element-wise ops on [B,32,4,H,W]) and
the tanh / mean restoration stay torch element-wise code.
"""
from __future__ import annotations

import os

from typing import List, Optional, Sequence

import torch
import torch.nn as nn

from . import functional as F_
from ._lib import ACT_LRELU, ACT_NONE, ACT_RELU, RehrError, device_of


# ----------------------------------------------------------------------------------------------------------------
# module tree (parameter containers with the reference's names; `forward`s run on the engine)
# ----------------------------------------------------------------------------------------------------------------
class identity(nn.Module):
    def __init__(self, *args, **kwargs):
        super().__init__()

    def forward(self, x):
        return x


class SEGating(nn.Module):
    """resnet_3D.py:100-116"""

    def __init__(self, inplanes, reduction=16):
        super().__init__()
        self.pool = nn.AdaptiveAvgPool3d(1)
        self.attn_layer = nn.Sequential(nn.Conv3d(inplanes, inplanes, kernel_size=1, stride=1, bias=True), nn.Sigmoid())


def _conv3d_simple(cin, cout, stride=1, padding=1):
    """Conv3DSimple (resnet_3D.py:19-33) with useBias = True (FLAVR_arch.py:133-134 flips the module global for n_outputs > 1)."""
    return nn.Conv3d(cin, cout, kernel_size=(3, 3, 3), stride=stride, padding=padding, bias=True)


class BasicBlock(nn.Module):
    """resnet_3D.py:118-151 (batchnorm = identity)"""
    expansion = 1

    def __init__(self, inplanes, planes, stride=1, downsample=None):
        super().__init__()
        self.conv1 = nn.Sequential(_conv3d_simple(inplanes, planes, stride), identity(), nn.ReLU(inplace=True))
        self.conv2 = nn.Sequential(_conv3d_simple(planes, planes), identity())
        self.fg = SEGating(planes)
        self.relu = nn.ReLU(inplace=True)
        self.downsample = downsample
        self.stride = stride


class VideoResNet(nn.Module):
    """resnet_3D.py:153-224 as instantiated by unet_18 (:238-261): BasicStem, 4 layers x 2 BasicBlocks, D never strided."""

    def __init__(self, img_channels=3, use_bias=True):
        super().__init__()
        if not use_bias:
            raise RehrError("FLAVR with n_outputs == 1 (bias-free encoder) is not used by the reference and not implemented")
        self.inplanes = 64
        self.stem = nn.Sequential(nn.Conv3d(img_channels, 64, kernel_size=(3, 7, 7), stride=(1, 2, 2), padding=(1, 3, 3), bias=True),
                                  identity(), nn.ReLU(inplace=False))
        self.layer1 = self._make_layer(64, 2, stride=1, temporal_stride=None)
        self.layer2 = self._make_layer(128, 2, stride=2, temporal_stride=1)
        self.layer3 = self._make_layer(256, 2, stride=2, temporal_stride=1)
        self.layer4 = self._make_layer(512, 2, stride=1, temporal_stride=1)
        for m in self.modules():  # resnet_3D.py:212-224
            if isinstance(m, nn.Conv3d):
                nn.init.kaiming_normal_(m.weight, mode='fan_out', nonlinearity='relu')
                if m.bias is not None:
                    nn.init.constant_(m.bias, 0)

    def _make_layer(self, planes, blocks, stride, temporal_stride):
        downsample = None
        if stride != 1 or self.inplanes != planes:
            ds_stride = (temporal_stride, stride, stride) if temporal_stride else (stride, stride, stride)
            downsample = nn.Sequential(nn.Conv3d(self.inplanes, planes, kernel_size=1, stride=ds_stride, bias=False), identity())
            stride = ds_stride
        layers = [BasicBlock(self.inplanes, planes, stride, downsample)]
        self.inplanes = planes
        for _ in range(1, blocks):
            layers.append(BasicBlock(self.inplanes, planes))
        return nn.Sequential(*layers)

    def forward(self, x):
        return encoder_forward(self, x)


class Conv_2d(nn.Module):
    """FLAVR_arch.py:24-38 (the reference passes bias=nn.InstanceNorm2d, which is merely truthy => bias=True, no norm)"""

    def __init__(self, in_ch, out_ch, kernel_size, stride=1, padding=0, bias=False, batchnorm=False):
        super().__init__()
        self.conv = nn.Sequential(nn.Conv2d(in_ch, out_ch, kernel_size=kernel_size, stride=stride, padding=padding, bias=bool(bias)))


class upConv3D(nn.Module):
    """FLAVR_arch.py:40-70, upmode="transpose" """

    def __init__(self, in_ch, out_ch, kernel_size, stride, padding, upmode="transpose", batchnorm=False):
        super().__init__()
        self.upmode = upmode
        self.upconv = nn.Sequential(nn.ConvTranspose3d(in_ch, out_ch, kernel_size=kernel_size, stride=stride, padding=padding),
                                    SEGating(out_ch))


class Conv_3d(nn.Module):
    """FLAVR_arch.py:72-88"""

    def __init__(self, in_ch, out_ch, kernel_size, stride=1, padding=0, bias=True, batchnorm=False):
        super().__init__()
        self.conv = nn.Sequential(nn.Conv3d(in_ch, out_ch, kernel_size=kernel_size, stride=stride, padding=padding, bias=bias),
                                  SEGating(out_ch))


class UNet_3D_3D(nn.Module):
    def __init__(self, img_channels, block, n_inputs, n_outputs, batchnorm=False, joinType="concat", upmode="transpose",
                 use_uncertainty=False):
        super().__init__()
        if block != "unet_18" or batchnorm or joinType != "concat" or upmode != "transpose":
            raise RehrError("only unet_18 / batchnorm=False / joinType='concat' / upmode='transpose' is implemented "
                            "(the configuration of train_all.py:336-345)")
        nf = [512, 256, 128, 64]
        self.out_channels = img_channels * n_outputs
        self.joinType, self.n_inputs, self.n_outputs = joinType, n_inputs, n_outputs
        self.img_channels, self.use_uncertainty = img_channels, use_uncertainty
        growth = 2
        self.lrelu = nn.LeakyReLU(0.2, True)
        self.encoder = VideoResNet(img_channels=img_channels, use_bias=n_outputs > 1)
        self.decoder = nn.Sequential(
            Conv_3d(nf[0], nf[1], kernel_size=3, padding=1, bias=True),
            upConv3D(nf[1] * growth, nf[2], kernel_size=(3, 4, 4), stride=(1, 2, 2), padding=(1, 1, 1)),
            upConv3D(nf[2] * growth, nf[3], kernel_size=(3, 4, 4), stride=(1, 2, 2), padding=(1, 1, 1)),
            Conv_3d(nf[3] * growth, nf[3], kernel_size=3, padding=1, bias=True),
            upConv3D(nf[3] * growth, nf[3], kernel_size=(3, 4, 4), stride=(1, 2, 2), padding=(1, 1, 1)))
        self.feature_fuse = Conv_2d(nf[3] * n_inputs, nf[3] * n_inputs if use_uncertainty else nf[3], kernel_size=3, stride=1,
                                    padding=1, bias=True)
        self.feature_fuse1 = Conv_2d(nf[3] * n_inputs, nf[3] * img_channels, kernel_size=1, stride=1, bias=True)
        self.tanh = torch.nn.Tanh()
        if self.use_uncertainty:
            self.uncertainty_early = Conv_2d(nf[3] * n_inputs, nf[3], kernel_size=1, stride=1, bias=True)
            self.softmax = nn.Softmax(dim=1)
            self.uncertainty_out = nn.Conv3d(nf[3] // n_outputs, 1, kernel_size=1, stride=1)
        self.outconv = nn.Sequential(nn.ReflectionPad2d(3), nn.Conv2d(nf[3], self.out_channels, kernel_size=7, stride=1, padding=0))

    def calc_out_patch_size(self, input_patch_size):
        """FLAVR_arch.py:158-167"""
        x = torch.rand(tuple([1, self.img_channels] + list(input_patch_size))).float()
        x = x.to(next(self.parameters()).device)
        with torch.no_grad():
            out = self(x)
        if self.use_uncertainty:
            out = out[0]
        patch_size = list(out.shape[2:])
        patch_size[0] *= self.n_inputs
        return patch_size

    def forward(self, images, return_inetermediate_uncertainty=False, return_inetermediate_feature=False):
        return flavr_forward(self, images, return_inetermediate_uncertainty, return_inetermediate_feature)


# ----------------------------------------------------------------------------------------------------------------
# engine forward (works on this module tree and on a reference-built one)
# ----------------------------------------------------------------------------------------------------------------
def _t3(v, fill=1):
    """kernel / stride (fill 1) or padding (fill 0) of a Conv3d, or of a Conv2d seen as a D=1 3-D op."""
    if isinstance(v, int):
        return (v, v, v)
    v = tuple(int(i) for i in v)
    return (fill,) + v if len(v) == 2 else v


def _w5(conv: nn.Module) -> torch.Tensor:
    """Conv2d weight [Co,Ci,kh,kw] seen as a D=1 3-D kernel."""
    return conv.weight.unsqueeze(2) if conv.weight.dim() == 4 else conv.weight


def _conv(conv, x, act=ACT_NONE, slope=0.0, want_pool=False, out_f32=False):
    two_d = conv.weight.dim() == 4
    k = ((1,) + tuple(conv.kernel_size)) if two_d else _t3(conv.kernel_size)
    s = ((1,) + tuple(conv.stride)) if two_d else _t3(conv.stride)
    p = ((0,) + tuple(conv.padding)) if two_d else _t3(conv.padding, 0)
    return F_.conv_act(x, _w5(conv), conv.bias, k, s, p, act=act, slope=slope, out_f32=out_f32, want_pool=want_pool)


def _gate(se: nn.Module, x, pool, residual=None, act=ACT_NONE, slope=0.0):
    a = se.attn_layer[0]
    return F_.se_gate(x, a.weight, a.bias, residual=residual, act=act, slope=slope, pool=pool)


def basic_block_forward(blk: nn.Module, x):
    """resnet_3D.py:140-151: relu(fg(conv2(relu(conv1(x)))) + downsample(x))"""
    out = _conv(blk.conv1[0], x, act=ACT_RELU)
    out, pool = _conv(blk.conv2[0], out, want_pool=True)
    residual = _conv(blk.downsample[0], x) if blk.downsample is not None else x
    return _gate(blk.fg, out, pool, residual=residual, act=ACT_RELU)


def encoder_forward(enc: nn.Module, images: torch.Tensor, n_feats: int = 5):
    """VideoResNet.forward (resnet_3D.py:183-189) -> 5 channels-last bf16 feature maps (`n_feats` < 5: only the first n_feats of
    them; the layers behind the last requested map are not run)."""
    stem = enc.stem[0]
    if stem.in_channels <= 4:
        x0 = F_.smallcin_conv_act(images, stem.weight, stem.bias, _t3(stem.kernel_size), _t3(stem.stride), _t3(stem.padding, 0),
                                  act=ACT_RELU)
    else:
        x0 = _conv(stem, F_.to_channels_last(images), act=ACT_RELU)
    feats = [x0]
    x = x0
    for layer in (enc.layer1, enc.layer2, enc.layer3, enc.layer4)[:max(0, int(n_feats) - 1)]:
        for blk in layer:
            x = basic_block_forward(blk, x)
        feats.append(x)
    return tuple(feats)


def _dec_conv(mod: nn.Module, x):
    """Conv_3d: conv -> SEGating, followed by the LeakyReLU(0.2) the caller applies (FLAVR_arch.py:187,...)."""
    y, pool = _conv(mod.conv[0], x, want_pool=True)
    return _gate(mod.conv[1], y, pool, act=ACT_LRELU, slope=0.2)


def _dec_up(mod: nn.Module, x):
    """upConv3D: ConvTranspose3d k(3,4,4) s(1,2,2) p(1,1,1) -> SEGating -> LeakyReLU(0.2)."""
    tc = mod.upconv[0]
    y = F_.conv_transpose(x, tc.weight, tc.bias, _t3(tc.kernel_size), _t3(tc.stride), _t3(tc.padding, 0))
    return _gate(mod.upconv[1], y, None, act=ACT_LRELU, slope=0.2)


_reflect_idx: dict = {}
FUSED_UASR = os.environ.get("REHR_FUSED_UASR", "1") != "0"   # the expert mixture of the UASR head as one kernel per direction


def _reflect_pad_hw(x: torch.Tensor, p: int) -> torch.Tensor:
    """nn.ReflectionPad2d(p) on a channels-last [B,1,H,W,C] tensor (index gathers; FLAVR_arch.py:153-156)."""
    h, w = x.shape[2], x.shape[3]
    key = (p, h, w, x.device)
    idx = _reflect_idx.get(key)
    if idx is None:      # built ON the device (no host copy: the forward stays capturable in a CUDA graph), once per shape
        def one(n):
            return torch.cat([torch.arange(p, 0, -1, device=x.device), torch.arange(n, device=x.device),
                              torch.arange(n - 2, n - 2 - p, -1, device=x.device)])
        idx = _reflect_idx[key] = (one(h), one(w))
    return x.index_select(2, idx[0]).index_select(3, idx[1])


class _UasrMixture(torch.autograd.Function):
    """The UASR head's expert mixture (FLAVR_arch.py:203-227,244-246) in one pass per direction: softmax over the 16 experts,
    img = sum p (tanh(o) + 1) / 2, seg = sum p o, uncertainty = sigmoid(sum p w + b) -- straight from the two channels-last fp32
    conv outputs (rehr_uasr_mixture_fwd / _bwd) instead of ~10 PyTorch passes over [B, 32, n_out, H, W] tensors."""

    @staticmethod
    def forward(ctx, out_cl, ue_cl, w, b, n_out):
        from . import functional as F_
        from ._lib import check, lib, ptr, stream_ptr
        bsz, _, h, wd, c = out_cl.shape
        experts = ue_cl.shape[4] // n_out
        if c != 2 * experts * n_out:
            raise RehrError("uasr mixture: feature_fuse1 must have twice the channels of uncertainty_early")
        out_cl, ue_cl = out_cl.contiguous(), ue_cl.contiguous()
        w32, b32 = w.detach().reshape(-1).float().contiguous(), b.detach().reshape(-1).float().contiguous()
        res = torch.empty((bsz, 2, n_out, h, wd), dtype=torch.float32, device=out_cl.device)
        unc = torch.empty((bsz, 1, n_out, h, wd), dtype=torch.float32, device=out_cl.device)
        check(lib().rehr_uasr_mixture_fwd(ptr(out_cl), ptr(ue_cl), ptr(w32), ptr(b32), ptr(res), ptr(unc), bsz, h * wd, n_out, experts,
                                          stream_ptr()), "uasr_mixture_fwd")
        F_._count()
        ctx.save_for_backward(out_cl, ue_cl, w32, b32)
        ctx.meta = (n_out, experts, w.shape, b.shape)
        return res, unc

    @staticmethod
    def backward(ctx, d_res, d_unc):
        from . import functional as F_
        from ._lib import check, lib, ptr, stream_ptr
        out_cl, ue_cl, w32, b32 = ctx.saved_tensors
        n_out, experts, wshape, bshape = ctx.meta
        bsz, _, h, wd, _ = out_cl.shape
        d_res = d_res.float().contiguous() if d_res is not None else None
        d_unc = d_unc.float().contiguous() if d_unc is not None else None
        d_out, d_ue = torch.empty_like(out_cl), torch.empty_like(ue_cl)
        blocks = lib().rehr_uasr_mixture_blocks(bsz * h * wd, n_out)
        partial = torch.empty((blocks, experts + 1), dtype=torch.float32, device=out_cl.device)
        check(lib().rehr_uasr_mixture_bwd(ptr(out_cl), ptr(ue_cl), ptr(w32), ptr(b32), ptr(d_res), ptr(d_unc), ptr(d_out), ptr(d_ue),
                                          ptr(partial), bsz, h * wd, n_out, experts, stream_ptr()), "uasr_mixture_bwd")
        F_._count()
        sums = partial.double().sum(0).float()
        return d_out, d_ue, sums[:experts].reshape(wshape), sums[experts:].reshape(bshape), None


def uasr_mixture(out_cl: torch.Tensor, ue_cl: torch.Tensor, uncertainty_out: nn.Module, n_outputs: int):
    """(res [B, 2, n_out, H, W], uncertainty [B, 1, n_out, H, W]) from the channels-last fp32 outputs [B, 1, H, W, C] of
    feature_fuse1 / uncertainty_early and the 1x1x1 `uncertainty_out` layer."""
    return _UasrMixture.apply(out_cl, ue_cl, uncertainty_out.weight, uncertainty_out.bias, int(n_outputs))


def flavr_forward(model: nn.Module, images: torch.Tensor, return_inetermediate_uncertainty=False,
                  return_inetermediate_feature=False):
    """UNet_3D_3D.forward (FLAVR_arch.py:169-248)."""
    if not images.is_cuda:
        raise RehrError("rehrseg_b200 runs on CUDA (sm_100a) only; there is no CPU path")
    with device_of(images):    # the library launches on the current device: follow the input's GPU
        return _flavr_forward(model, images, return_inetermediate_uncertainty, return_inetermediate_feature)


def _flavr_forward(model: nn.Module, images: torch.Tensor, return_inetermediate_uncertainty=False,
                   return_inetermediate_feature=False):
    # in-place on the caller's tensor, exactly like the reference (callers clone first: train_all.py:98, sr_utils.py:125)
    mean_ = images[:, 0:1, ...].mean(2, keepdim=True).mean(3, keepdim=True).mean(4, keepdim=True)
    images[:, 0:1, ...] = images[:, 0:1, ...] - mean_

    x_0, x_1, x_2, x_3, x_4 = encoder_forward(model.encoder, images)
    if return_inetermediate_feature:
        return tuple(F_.from_channels_last(t) for t in (x_0, x_1, x_2, x_3, x_4))

    dec = model.decoder
    dx_3 = torch.cat((_dec_conv(dec[0], x_4), x_3), dim=4)
    dx_2 = torch.cat((_dec_up(dec[1], dx_3), x_2), dim=4)
    dx_1 = torch.cat((_dec_up(dec[2], dx_2), x_1), dim=4)
    dx_0 = torch.cat((_dec_conv(dec[3], dx_1), x_0), dim=4)
    dx_out = _dec_up(dec[4], dx_0)                                 # [B, D, H, W, 64]
    b, d, h, w, c = dx_out.shape
    # torch.cat(torch.unbind(dx_out, 2), 1): channel index = depth * 64 + c
    dx_out = dx_out.permute(0, 2, 3, 1, 4).reshape(b, 1, h, w, d * c)

    def nchw(t):  # [B,1,H,W,C] fp32 -> [B,C,H,W]
        return t[:, 0].permute(0, 3, 1, 2)

    if model.use_uncertainty:
        fused = _conv(model.feature_fuse.conv[0], dx_out, act=ACT_LRELU, slope=0.2)
        if FUSED_UASR and not return_inetermediate_uncertainty:
            out_cl = _conv(model.feature_fuse1.conv[0], fused, out_f32=True)           # [B, 1, H, W, 32 * n_out]
            ue_cl = _conv(model.uncertainty_early.conv[0], fused, out_f32=True)        # [B, 1, H, W, 16 * n_out]
            if ue_cl.shape[4] == 16 * model.n_outputs and out_cl.shape[4] == 32 * model.n_outputs:
                return uasr_mixture(out_cl, ue_cl, model.uncertainty_out, model.n_outputs)
        out = nchw(_conv(model.feature_fuse1.conv[0], fused, out_f32=True))
        out = torch.stack(torch.split(out, out.shape[1] // model.n_outputs, dim=1), dim=2)          # [B, 32, n_out, H, W]
        ue = nchw(_conv(model.uncertainty_early.conv[0], fused, out_f32=True))
        ue = torch.stack(torch.split(ue, ue.shape[1] // model.n_outputs, dim=1), dim=2)              # [B, 16, n_out, H, W]
        usm = torch.softmax(ue, dim=1)
        if return_inetermediate_uncertainty:
            ne = usm.shape[1]
            imgs = [(torch.tanh(out[:, 2 * i:2 * i + 1]) + 1) / 2 for i in range(ne)]
            return imgs, [usm[:, i:i + 1] for i in range(ne)], [out[:, 2 * i + 1:2 * i + 2] for i in range(ne)]
        img = ((torch.tanh(out[:, 0::2]) + 1) / 2 * usm).sum(1, keepdim=True)
        seg = (out[:, 1::2] * usm).sum(1, keepdim=True)
        res = torch.cat([img, seg], dim=1)
        uo = model.uncertainty_out
        unc = torch.sigmoid(torch.einsum('bcdhw,c->bdhw', usm, uo.weight.reshape(-1)).unsqueeze(1) + uo.bias.reshape(1, 1, 1, 1, 1))
        return res, unc

    fused = _conv(model.feature_fuse.conv[0], dx_out, act=ACT_LRELU, slope=0.2)
    oc = model.outconv[1]
    pad = model.outconv[0].padding[0] if hasattr(model.outconv[0], "padding") else 3
    out = nchw(_conv(oc, _reflect_pad_hw(fused, int(pad)), out_f32=True))                              # [B, 8, H, W]
    outs = torch.split(out, model.img_channels, dim=1)
    m2 = mean_.squeeze(2)
    if model.img_channels > 1:
        outs = [torch.cat([torch.tanh(o[:, 0:1] + m2), o[:, 1:2]], dim=1) for o in outs]
    else:
        outs = [o + m2 for o in outs]
    return torch.stack(outs, dim=2)


class _EngineForward:
    def forward(self, images, return_inetermediate_uncertainty=False, return_inetermediate_feature=False):
        return flavr_forward(self, images, return_inetermediate_uncertainty, return_inetermediate_feature)


def convert(model: nn.Module) -> nn.Module:
    """Route a reference-built UNet_3D_3D through the engine in place (parameters / state_dict untouched)."""
    for attr in ("encoder", "decoder", "feature_fuse", "outconv", "n_outputs", "use_uncertainty"):
        if not hasattr(model, attr):
            raise RehrError(f"convert(): {type(model).__name__} has no .{attr}; expected the reference UNet_3D_3D")
    if getattr(model, "joinType", "concat") != "concat":
        raise RehrError("only joinType='concat' is implemented")
    if isinstance(model, _EngineForward):
        return model
    cls = model.__class__
    model.__class__ = type("B200" + cls.__name__, (_EngineForward, cls), {})
    return model


# ----------------------------------------------------------------------------------------------------------------
# window sweeps
# ----------------------------------------------------------------------------------------------------------------
def _windows(vol: torch.Tensor, dim: int) -> torch.Tensor:
    """All Z-1 four-slice windows of `vol` along `dim` with the reference's zero padding at both ends
    (sr_utils.py:115-123, train_all.py:89-97): window st covers slices st-1 .. st+2.  Returns a new leading window axis."""
    z = vol.shape[dim]
    pad_shape = list(vol.shape)
    pad_shape[dim] = 1
    zero = vol.new_zeros(pad_shape)
    padded = torch.cat([zero, vol, zero], dim=dim)             # slice s of vol sits at s + 1
    return torch.stack([padded.narrow(dim, st, 4) for st in range(z - 1)], dim=0)


def apply_to_vol_flavr(model, image: torch.Tensor, pred_out_idx=None, max_batch: int = 8, use_graph: bool = True) -> torch.Tensor:
    """utils/sr_utils.py:102-135.  `image` [Z, C, X, Y] -> [4(Z-1), C', Y, X] (the reference's axis order), computed with
    `max_batch` windows per forward instead of one, on the device (the reference moves every window result to the CPU).
    `use_graph`: the full-size window batches all have one shape, so their forward (~200 launches of 5-300 us, bound by the host
    when issued from Python) is captured once and replayed (graphs.graphed; re-captured when the weights change)."""
    if image.shape[0] < 3:
        raise RehrError("apply_to_vol_flavr needs at least 3 slices (the reference indexes image[0:3])")
    dev = image.device if image.is_cuda else torch.device("cuda", torch.cuda.current_device())
    image = image.to(dev)
    ox, oy = image.shape[2], image.shape[3]
    px, py = (-ox) % 16, (-oy) % 16
    if px or py:
        image = torch.nn.functional.pad(image, (0, py, 0, px))
    win = _windows(image, 0)                                       # [Z-1, 4, C, X, Y]
    win = win.permute(0, 2, 1, 4, 3)                               # -> [Z-1, C, 4, Y, X]   (batch.permute(1,0,3,2) per window)
    outs = []
    with torch.no_grad():
        for s in range(0, win.shape[0], max_batch):
            batch = win[s:s + max_batch].contiguous()
            if use_graph and batch.shape[0] == max_batch and win.shape[0] >= 2 * max_batch:
                from .graphs import graphed
                sr = graphed(model, batch)(batch)        # static outputs: copied out below before the next replay
                replayed = True
            else:
                sr = model(batch.clone())                 # (the forward subtracts the mean from its input in place)
                replayed = False
            if pred_out_idx is not None and isinstance(sr, tuple):
                sr = sr[pred_out_idx]
            sr = sr[:, :, :, :oy, :ox]
            outs.append(sr.clone() if replayed else sr)
    res = torch.cat(outs, dim=0)                                   # [Z-1, C', 4, Y, X]
    return res.permute(0, 2, 1, 3, 4).reshape(-1, res.shape[1], res.shape[3], res.shape[4])


def zscore_normalization(image: torch.Tensor) -> torch.Tensor:
    """utils/seg_utils.py:137-148 (tensor branch): channel 0 of every sample is z-scored IN PLACE through a view, so the
    caller's tensor changes (the stage-2 student then sees the re-normalised image, train_all.py:86,533-534)."""
    outs = []
    for i in range(image.shape[0]):
        img = image[i:i + 1, 0, ...]
        mean = img.mean()
        std = img.std()
        img -= mean
        img /= torch.clamp(std, min=1e-8)
        outs.append(img)
    return torch.stack(outs, dim=0)


def get_intermediate_features(model_sr, img_lr: torch.Tensor, label_lr: torch.Tensor, device=None, normalize=None,
                              max_batch: int = 8, keys: Optional[Sequence[int]] = None) -> dict:
    """train_all.py:85-112: teacher encoder features of every 4-slice window, slice 1 of each window (and slice 2 of the
    last) stitched along D.  `normalize` = the reference's `zscore_normalization` (utils/seg_utils.py:137-148), applied in
    place to `img_lr` exactly as the reference does; the D-1 windows are batched through the encoder.
    `keys` (opt-in, not a reference argument): the feature maps the caller will read -- the stage-2 loop reads only `[1]`
    (train_all.py:550).  None = all five, as the reference returns them; with keys an engine teacher stops after the last
    requested map and converts only those (identical values for the returned entries)."""
    if normalize is not None:
        img_lr = normalize(img_lr)
    inp = torch.cat((img_lr, label_lr), dim=1)                     # [B, 2, D, H, W]
    b = inp.shape[0]
    win = _windows(inp, 2)                                          # [D-1, B, 2, 4, H, W]
    nwin = win.shape[0]
    flat = win.reshape(nwin * b, *win.shape[2:])
    if isinstance(model_sr, (UNet_3D_3D, _EngineForward)) and flat.is_cuda:
        return _stitched_features_cl(model_sr, flat, nwin, b, max_batch, keys)
    chunks: Optional[List[List[torch.Tensor]]] = None
    for s in range(0, flat.shape[0], max_batch * b):
        feats = model_sr(flat[s:s + max_batch * b].clone(), return_inetermediate_feature=True)
        if chunks is None:
            chunks = [[] for _ in feats]
        for i, f in enumerate(feats):
            chunks[i].append(f)
    out = {}
    for i, parts in enumerate(chunks):
        f = torch.cat(parts, dim=0)                                 # [(D-1)*B, C, 4, h, w]
        f = f.reshape(nwin, b, *f.shape[1:])
        mid = f[:, :, :, 1]                                         # slice 1 of every window  [D-1, B, C, h, w]
        last = f[-1:, :, :, 2]                                      # slice 2 of the last window
        out[i] = torch.cat([mid, last], dim=0).permute(1, 2, 0, 3, 4).contiguous()   # [B, C, D, h, w]
    if keys is not None:
        out = {int(k): out[int(k)] for k in keys}
    return out


def _stitched_features_cl(model_sr, flat: torch.Tensor, nwin: int, b: int, max_batch: int, keys: Optional[Sequence[int]] = None) -> dict:
    """Engine teacher: the loop above keeps slice 1 of every window (and slice 2 of the last) of each feature map, so only those
    depth slices of the channels-last 16-bit encoder outputs are converted to NCDHW fp32 -- a quarter of the adapter traffic of
    converting all four slices first (2.4 GB per C4 step).  Same values: the conversion is element-wise."""
    want = sorted({int(k) for k in keys}) if keys is not None else list(range(5))
    if not want or want[0] < 0 or want[-1] > 4:
        raise RehrError("get_intermediate_features: keys must be a non-empty subset of 0..4")
    mids: dict = {k: [] for k in want}
    lasts: dict = {}
    total = flat.shape[0]
    for s in range(0, total, max_batch * b):
        images = flat[s:s + max_batch * b].clone()
        with device_of(images):
            # the head of UNet_3D_3D.forward (FLAVR_arch.py:171-175): mean of channel 0 removed in place, then the encoder
            mean_ = images[:, 0:1, ...].mean(2, keepdim=True).mean(3, keepdim=True).mean(4, keepdim=True)
            images[:, 0:1, ...] = images[:, 0:1, ...] - mean_
            feats = encoder_forward(model_sr.encoder, images, n_feats=want[-1] + 1)   # channels-last [n, 4, h, w, C]
            def plane(f, lo, k):   # depth slice k of samples lo: as NCHW fp32 (a 16-bit payload mark does not survive slicing)
                if F_.is_h(f):
                    return F_.from_channels_last(f)[lo:, :, k]
                return F_.from_channels_last(f[lo:, k:k + 1].contiguous())[:, :, 0]
            for i in want:
                mids[i].append(plane(feats[i], 0, 1))                   # [n, C, h, w] fp32
            if s + images.shape[0] == total:                            # this chunk ends with the last window's b samples
                lasts = {i: plane(feats[i], feats[i].shape[0] - b, 2) for i in want}
    out = {}
    for i, parts in mids.items():
        mid = torch.cat(parts, dim=0).reshape(nwin, b, *parts[0].shape[1:])                  # [D-1, B, C, h, w]
        out[i] = torch.cat([mid, lasts[i][None]], dim=0).permute(1, 2, 0, 3, 4).contiguous()  # [B, C, D, h, w]
    return out


def sr_volume_orientations(model, image: torch.Tensor, angles=(0,), pred_out_idx=0, fuse="mean", group=None, max_batch: int = 8):
    """The tensor part of `inference_flavr` (utils/sr_utils.py:157-175): for every in-plane angle rotate `image`
    [hr, hr, lr, C] with rotate_vol_2d, move the axes to (hr, C, lr, hr), sweep the 4-slice windows through the network
    (apply_to_vol_flavr), move the axes back, undo the rotation, then fuse the orientations -- "mean" (what the reference
    runs, torch.mean(torch.stack(...))) or "fba_inf" / "fba_<p>" (utils/fba.py).  The reference ships angles = [0]; the SMORE
    lineage uses [0, 90].  With a torch.distributed `group` the angles are dealt round-robin over the ranks (one process per
    GPU) and the per-orientation volumes are all-gathered before the (replicated) fusion."""
    from . import volume_ops as vo
    import torch.distributed as dist
    rank = dist.get_rank(group) if (group is not None or dist.is_initialized()) and dist.is_initialized() else 0
    world = dist.get_world_size(group) if dist.is_initialized() else 1
    dev = image.device if image.is_cuda else torch.device("cuda", torch.cuda.current_device())
    image = image.to(dev)
    mine = {}
    for i, angle in enumerate(angles):
        if i % world != rank:
            continue
        rot = vo.rotate_vol_2d(image, angle).permute(0, 3, 2, 1)
        res = apply_to_vol_flavr(model, rot, pred_out_idx, max_batch=max_batch).permute(0, 3, 1, 2)
        mine[i] = vo.rotate_vol_2d(res.contiguous(), -angle)
    if world > 1:
        shape = None
        for v in mine.values():
            shape = v.shape
        # every orientation comes back in the un-rotated frame, so all ranks hold equally shaped volumes
        shp = torch.tensor(list(shape) if shape is not None else [0, 0, 0, 0], device=dev)
        dist.all_reduce(shp, op=dist.ReduceOp.MAX, group=group)
        preds = []
        for i in range(len(angles)):
            buf = mine[i].contiguous() if i in mine else torch.empty(tuple(int(s) for s in shp), dtype=torch.float32, device=dev)
            dist.broadcast(buf, src=dist.get_global_rank(group, i % world) if group is not None else i % world, group=group)
            preds.append(buf)
    else:
        preds = [mine[i] for i in range(len(angles))]
    if fuse == "mean":
        return vo.mean_fuse(preds)
    if fuse.startswith("fba"):
        p = fuse.split("_", 1)[1] if "_" in fuse else "infinity"
        return vo.fba(preds, "infinity" if p in ("inf", "infinity") else float(p))
    raise RehrError(f"unknown fusion {fuse!r}")
