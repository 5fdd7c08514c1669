"""B200-native WDSR (the SMORE 2-D stage-one super-resolution network), drop-in for the reference's models/wdsr.py:58-95.

Same constructor `WDSR(out_channel, n_resblocks, num_channels, scale)`, same module tree (`head`, `body.N.body.{0,2,3}`,
`tail.conv0`, `skip.conv0`, every Conv2d weight-normalised) and therefore the same `state_dict` keys (`*.weight_g`, `*.weight_v`,
`*.bias`), the same `calc_out_patch_size`, and the reference's `train_sr` loop (train_all.py:300-306) runs it unchanged.  The
forward runs the 2-D convolutions on the 3-D conv engine with depth 1:

    head  Conv2d(out_channel -> n, 3x3)                  small-Cin direct conv from the caller's NCHW fp32 (functional.smallcin_conv_act)
    block Conv2d(n -> 4n, 1x1) + ReLU, (4n -> 0.8n, 1x1), (0.8n -> n, 3x3), + x   tcgen05 implicit GEMMs (functional.conv_act); the
          0.8n = 25-channel bottleneck is zero-padded to 32 channels (zero filters / zero weights: exact)
    tail  Conv2d(n -> scale*out_channel, 3x3), skip Conv2d(out_channel -> scale*out_channel, 5x5), pixel shuffle along x, sum

The weight normalisation w = g * v / ||v|| is a handful of tiny PyTorch ops on the parameters (autograd carries the weight
gradient back to g and v).  `resize(x, (1 / scale0, 1), order=3)` (models/wdsr.py:87) is a third-party cubic resampler whose
source is unavailable; for integer `scale` it resamples by the factor 1 -- the identity -- and that is
what is implemented; a rational scale raises.
"""
from __future__ import annotations

import torch
import torch.nn as nn
import torch.nn.functional as F

from . import functional as F_
from ._lib import ACT_NONE, ACT_RELU, C, RehrError, check, device_of, lib, ptr, rt, stream_ptr


def pixel_shuffle(x: torch.Tensor, scale: int) -> torch.Tensor:
    """models/wdsr.py:13-21: [B, C*scale, X, Y] -> [B, C, X*scale, Y] (channel c*scale + s becomes row x*scale + s)."""
    b, c, nx, ny = x.shape
    c //= scale
    return x.contiguous().view(b, c, scale, nx, ny).permute(0, 1, 3, 2, 4).contiguous().view(b, c, nx * scale, ny)


class Upsample(nn.Module):
    def __init__(self, out_channel, num_channels, scale, kernel_size, wn):
        super().__init__()
        self.scale = scale
        self.conv0 = wn(nn.Conv2d(num_channels, scale * out_channel, kernel_size, padding=(kernel_size - 1) // 2))


class Block(nn.Module):
    def __init__(self, n_feats, wn, act=None, res_scale=1):
        super().__init__()
        self.res_scale = res_scale
        expand, linear = 4, 0.8
        self.body = nn.Sequential(wn(nn.Conv2d(n_feats, n_feats * expand, 1, padding=0)), act if act is not None else nn.ReLU(True),
                                  wn(nn.Conv2d(n_feats * expand, int(n_feats * linear), 1, padding=0)),
                                  wn(nn.Conv2d(int(n_feats * linear), n_feats, 3, padding=1)))


def _w(conv: nn.Module) -> torch.Tensor:
    """Effective weight of a weight-normalised Conv2d as a [Cout, Cin, 1, kh, kw] tensor (g * v / ||v||, norm over all but dim 0)."""
    if hasattr(conv, "weight_g"):
        w = torch._weight_norm(conv.weight_v, conv.weight_g, 0)
    else:
        w = conv.weight
    return w.unsqueeze(2)


def _pad_dim(t: torch.Tensor, dim: int, to: int) -> torch.Tensor:
    extra = to - t.shape[dim]
    if extra <= 0:
        return t
    pad = [0, 0] * (t.dim() - dim - 1) + [0, extra]
    return F.pad(t, pad)


class _AddScaled(torch.autograd.Function):
    """res * scale + x on channels-last bf16 tensors (the residual connection of a WDSR block, models/wdsr.py:53-55)."""

    @staticmethod
    def forward(ctx, res, x, scale):
        res, x = F_.as_cl(res), F_.as_cl(x)
        n, c = res.shape[0], res.shape[4]
        gate = torch.full((n, c), float(scale), dtype=torch.float32, device=res.device)
        y = torch.empty_like(res)
        rt_, xt, yt = rt(res), rt(x), rt(y)
        check(lib().rehr_segate_scale_add_act(C.byref(rt_), ptr(gate), C.byref(xt), ACT_NONE, 0.0, C.byref(yt), stream_ptr()),
              "scale_add")
        F_._count()
        ctx.scale = float(scale)
        return y

    @staticmethod
    def backward(ctx, dy):
        return (dy if ctx.scale == 1.0 else dy * ctx.scale), dy, None


def _pad8(n: int) -> int:
    return (n + 7) // 8 * 8


def block_forward(blk: Block, h: torch.Tensor) -> torch.Tensor:
    c0, c2, c3 = blk.body[0], blk.body[2], blk.body[3]
    k1, s1, p0 = (1, 1, 1), (1, 1, 1), (0, 0, 0)
    t = F_.conv_act(h, _w(c0), c0.bias, k1, s1, p0, act=ACT_RELU)
    mid = _pad8(c2.out_channels)          # 0.8 * n_feats (25 for n = 32) -> channel counts the 16-byte tile rows need
    mid = (mid + 15) // 16 * 16
    t = F_.conv_act(t, _pad_dim(_w(c2), 0, mid), _pad_dim(c2.bias, 0, mid), k1, s1, p0)
    t = F_.conv_act(t, _pad_dim(_w(c3), 1, mid), c3.bias, (1, 3, 3), s1, (0, 1, 1))
    return _AddScaled.apply(t, h, blk.res_scale)


class WDSR(nn.Module):
    def __init__(self, out_channel, n_resblocks, num_channels, scale):
        super().__init__()
        self._scale1 = int(scale)
        self._scale0 = scale / float(self._scale1)
        self.out_channel = out_channel
        wn = lambda m: torch.nn.utils.weight_norm(m)  # noqa: E731  (same (deprecated) API as the reference -> weight_g / weight_v keys)
        act = nn.ReLU(True)
        self.head = wn(nn.Conv2d(out_channel, num_channels, 3, padding=1))
        self.body = nn.Sequential(*[Block(num_channels, act=act, res_scale=1, wn=wn) for _ in range(n_resblocks)])
        self.tail = Upsample(out_channel, num_channels, self._scale1, 3, wn)
        self.skip = Upsample(out_channel, out_channel, self._scale1, 5, wn)

    def calc_out_patch_size(self, input_patch_size):
        x = torch.rand([1, self.out_channel] + list(input_patch_size)).float().to(next(self.parameters()).device)
        with torch.no_grad():
            return list(self(x).shape[2:])

    def forward(self, x: torch.Tensor) -> torch.Tensor:
        return wdsr_forward(self, x)


def wdsr_forward(model: nn.Module, x: torch.Tensor) -> torch.Tensor:
    """WDSR.forward (models/wdsr.py:86-95) on the engine; x [B, out_channel, X, Y] fp32 NCHW -> [B, out_channel, X*scale, Y] fp32."""
    if not x.is_cuda:
        raise RehrError("rehrseg_b200 runs on CUDA (sm_100a) only; there is no CPU path")
    if abs(model._scale0 - 1.0) > 1e-12:
        raise RehrError("WDSR with a rational scale needs the third-party cubic `resize` (models/wdsr.py:87), whose source is not "
                        "available; integer scales (resize by 1 = identity) are implemented")
    if model.out_channel > 4:
        raise RehrError("WDSR head / skip convolutions use the small-Cin direct kernel (out_channel <= 4; the reference uses 2)")
    with device_of(x):
        x5 = x.float().unsqueeze(2)                              # NCDHW with D = 1
        one = (1, 1, 1)
        s = F_.smallcin_conv_act(x5, _w(model.skip.conv0), model.skip.conv0.bias, (1, 5, 5), one, (0, 2, 2))
        h = F_.smallcin_conv_act(x5, _w(model.head), model.head.bias, (1, 3, 3), one, (0, 1, 1))
        for blk in model.body:
            h = block_forward(blk, h)
        t = F_.conv_act(h, _w(model.tail.conv0), model.tail.conv0.bias, (1, 3, 3), one, (0, 1, 1))
        t = F_.from_channels_last(t).squeeze(2)                   # [B, scale*out_channel, X, Y] fp32
        s = F_.from_channels_last(s).squeeze(2)
        return pixel_shuffle(t, model._scale1) + pixel_shuffle(s, model._scale1)


class _EngineForward:
    def forward(self, x):
        return wdsr_forward(self, x)


def convert(model: nn.Module) -> nn.Module:
    """Route an already-built reference WDSR through the engine in place (parameters and state_dict untouched)."""
    for attr in ("head", "body", "tail", "skip", "_scale0", "_scale1", "out_channel"):
        if not hasattr(model, attr):
            raise RehrError(f"convert(): {type(model).__name__} has no .{attr}; expected the reference WDSR")
    if isinstance(model, _EngineForward):
        return model
    cls = model.__class__
    model.__class__ = type("B200" + cls.__name__, (_EngineForward, cls), {})
    return model


def apply_to_vol_smore(model, image: torch.Tensor, batch_size: int) -> torch.Tensor:
    """utils/sr_utils.py:20-31: slabs of `batch_size` slices, last two axes swapped, through the model; results gathered on the host
    like the reference (`.detach().cpu()` per slab)."""
    result = []
    for st in range(0, image.shape[0], batch_size):
        batch = image[st:st + batch_size].permute(0, 1, 3, 2)
        with torch.inference_mode():
            result.append(model(batch.contiguous()).detach().cpu())
    return torch.cat(result, dim=0)
