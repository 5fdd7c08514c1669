"""CPU restatement (plain PyTorch fp32) of the reference's WDSR network -- TEST INFRASTRUCTURE.

Follows models/wdsr.py: pixel_shuffle :13-21, Upsample :24-35, Block :38-56 (1x1 expand x4 + ReLU, 1x1 to int(0.8 n), 3x3, residual),
WDSR :58-95 (weight-normalised Conv2d everywhere, `skip` 5x5 + `tail` 3x3 both pixel-shuffled along x and summed).
`resize(x, (1 / scale0, 1), order=3)` (:87) is a third-party cubic resampler that is not available (PARITY UNPINNED); for integer
scales its factor is 1 and it is taken to be the identity, which is also how tests/golden/wdsr_small.npz was produced from the
reference's own module (oracle/make_golden.py::wdsr_fixture patches `resize` with the identity).
"""
from __future__ import annotations

import torch
import torch.nn.functional as F
from torch import nn


def pixel_shuffle(x, scale):
    b, c, nx, ny = x.shape
    c = c // scale
    return x.contiguous().view(b, c, scale, nx, ny).permute(0, 1, 3, 2, 4).contiguous().view(b, c, nx * scale, ny)


def _wn(m):
    return torch.nn.utils.weight_norm(m)


class RefUpsample(nn.Module):
    def __init__(self, out_channel, num_channels, scale, kernel_size):
        super().__init__()
        self.scale = scale
        self.conv0 = _wn(nn.Conv2d(num_channels, scale * out_channel, kernel_size, padding=(kernel_size - 1) // 2))

    def forward(self, x):
        return pixel_shuffle(self.conv0(x), self.scale)


class RefBlock(nn.Module):
    def __init__(self, n_feats, res_scale=1):
        super().__init__()
        self.res_scale = res_scale
        self.body = nn.Sequential(_wn(nn.Conv2d(n_feats, n_feats * 4, 1, padding=0)), nn.ReLU(True),
                                  _wn(nn.Conv2d(n_feats * 4, int(n_feats * 0.8), 1, padding=0)),
                                  _wn(nn.Conv2d(int(n_feats * 0.8), n_feats, 3, padding=1)))

    def forward(self, x):
        return self.body(x) * self.res_scale + x


class RefWDSR(nn.Module):
    def __init__(self, out_channel, n_resblocks, num_channels, scale):
        super().__init__()
        self._scale1 = int(scale)
        self._scale0 = scale / float(self._scale1)
        self.out_channel = out_channel
        self.head = _wn(nn.Conv2d(out_channel, num_channels, 3, padding=1))
        self.body = nn.Sequential(*[RefBlock(num_channels) for _ in range(n_resblocks)])
        self.tail = RefUpsample(out_channel, num_channels, self._scale1, 3)
        self.skip = RefUpsample(out_channel, out_channel, self._scale1, 5)

    def forward(self, x):
        if abs(self._scale0 - 1.0) > 1e-12:
            raise NotImplementedError("rational scale: the third-party `resize` is unavailable")
        s = self.skip(x)
        return self.tail(self.body(self.head(x))) + s


def build(out_channel=2, n_resblocks=2, num_channels=32, scale=4.0, seed=1234) -> RefWDSR:
    torch.manual_seed(seed)
    return RefWDSR(out_channel, n_resblocks, num_channels, scale)
