"""CPU oracle of the SR-stage sample synthesis: the tensor part of
`TrainSetMultiple.__getitem__` (utils/train_set.py:337-434) and the blur pre-filter of `load_img` (:321-333).
TEST INFRASTRUCTURE: imported only by tests/, __graft_entry__.smoke() and oracle/make_golden.py.

Third-party piece: `resize.pytorch.resize` @ gitlab iacl/resize 7edd72d (requirements.txt:80) -- its source is not in the
image and not in /root/reference, so `resize_standin` below DEFINES the resampling this repo uses in its place ("parity
unpinned" for that one function):
    output length n_out = round(n_in / d) per axis; output sample i sits at p = (i + 0.5) * d - 0.5 input samples (same field of
    view); order 3 = cubic convolution with A = -0.75 (the kernel of torch's bicubic grid_sample, which resize.pytorch is
    documented to call) over floor(p) - 1 .. floor(p) + 2 with indices clamped to the volume; order 0 = nearest, floor(p + 0.5).
Everything AROUND it is pinned: oracle/make_golden.py runs the reference's own `__getitem__` with this stand-in injected as
`resize.pytorch.resize` and seeded `random`, and `train_sample` below must reproduce its outputs bit for bit
(tests/test_oracle_golden.py)."""
from __future__ import annotations

import random as _random

import numpy as np
import torch

from .volume import get_pads


def _resample_axis(x: torch.Tensor, axis: int, d: float, order: int) -> torch.Tensor:
    n_in = x.shape[axis]
    n_out = int(round(n_in / d))
    i = torch.arange(n_out, dtype=torch.float32)
    p = (i + 0.5) * np.float32(d) - 0.5
    xm = x.movedim(axis, -1).float()
    if order == 0:
        j = torch.floor(p + 0.5).long().clamp(0, n_in - 1)
        out = xm[..., j]
    elif order == 3:
        fl = torch.floor(p)
        f = p - fl
        j0 = fl.long()
        A = np.float32(-0.75)
        g = 1.0 - f
        w = [((A * (f + 1) - 5 * A) * (f + 1) + 8 * A) * (f + 1) - 4 * A,
             ((A + 2) * f - (A + 3)) * f * f + 1,
             ((A + 2) * g - (A + 3)) * g * g + 1,
             ((A * (g + 1) - 5 * A) * (g + 1) + 8 * A) * (g + 1) - 4 * A]
        out = None
        for k in range(4):
            idx = (j0 - 1 + k).clamp(0, n_in - 1)
            term = w[k] * xm[..., idx]
            out = term if out is None else out + term
    else:
        raise NotImplementedError(f"resize stand-in: order {order}")
    return out.movedim(-1, axis).contiguous()


def resize_standin(x: torch.Tensor, dxy, order: int = 3) -> torch.Tensor:
    """Stand-in of `resize.pytorch.resize(image[B, C, X, Y], (dx, dy), order=...)` as defined in the module docstring."""
    if x.dim() != 4 or len(dxy) != 2:
        raise NotImplementedError("resize stand-in: [B, C, X, Y] tensors and two step factors")
    out = x
    for axis, d in ((2, float(dxy[0])), (3, float(dxy[1]))):
        if d != 1.0:
            out = _resample_axis(out, axis, d, order)
    return out.float() if out is x else out


def blur_prefilter(image_xyzc: np.ndarray, kernel: torch.Tensor):
    """load_img, utils/train_set.py:321-333: (image_x_rgb [Z,1,X,Y] blurred along X, image_y_rgb [Z,1,Y,X] blurred along Y)."""
    import torch.nn.functional as F
    image_x = torch.from_numpy(image_xyzc.transpose(2, 3, 0, 1))[:, 0:1]
    image_y = torch.from_numpy(image_xyzc.transpose(2, 3, 1, 0))[:, 0:1]
    return (F.conv2d(image_x, kernel, padding="same").numpy(), F.conv2d(image_y, kernel, padding="same").numpy())


def _target_pad(img: np.ndarray, target_dims):
    pads = tuple(get_pads(t, d) for t, d in zip(target_dims, img.shape))
    return np.pad(img, pads, mode="constant")


def train_sample(img_hr: np.ndarray, label_hr: np.ndarray, img_filtered_x, img_filtered_y, patch_size, slice_separation: float,
                 blur: bool = True, random_flip: bool = True, rng=_random):
    """`TrainSetMultiple.__getitem__` with `train_transform=None` (utils/train_set.py:337-434), statement by statement, drawing
    from `rng` in the reference's order.  img_hr / label_hr: [X, Y, Z, 1]; filtered volumes as `blur_prefilter` returns them."""
    if blur:
        if rng.random() < 0.5:
            img_hr = np.transpose(img_hr, (1, 0, 2, 3))
            label_hr = np.transpose(label_hr, (1, 0, 2, 3))
            img_lr = img_filtered_y
        else:
            img_lr = img_filtered_x
    else:
        if rng.random() < 0.5:
            img_hr = np.transpose(img_hr, (1, 0, 2, 3))
            label_hr = np.transpose(label_hr, (1, 0, 2, 3))
    ps = patch_size
    x_0 = rng.randint(0, max(img_hr.shape[0] - ps[0], 0))
    y_0 = rng.randint(0, max(img_hr.shape[1] - ps[1], 0))
    z_0 = rng.randint(0, max(img_hr.shape[2] - ps[2], 0))
    img_hr = img_hr[x_0:x_0 + ps[0], y_0:y_0 + ps[1], z_0:z_0 + ps[2], :]
    patch_label_hr = label_hr[x_0:x_0 + ps[0], y_0:y_0 + ps[1], z_0:z_0 + ps[2], :].astype("float32")
    img_hr = img_hr.transpose(2, 3, 0, 1)
    patch_label_hr = patch_label_hr.transpose(2, 3, 0, 1)
    target_shape = [max(s, p) for s, p in zip(img_hr.shape, (ps[2], 1, ps[0], ps[0]))]      # (ps[0] twice: the reference's line 362)
    img_hr = torch.from_numpy(_target_pad(img_hr, target_shape))
    patch_label_hr = torch.from_numpy(_target_pad(patch_label_hr, target_shape))
    if blur:
        img_lr = img_lr[z_0:z_0 + ps[2], :, x_0:x_0 + ps[0], y_0:y_0 + ps[1]]
        img_lr = torch.from_numpy(_target_pad(img_lr, target_shape))
    else:
        img_lr = img_hr.detach()
    img_hr = torch.cat((img_hr, patch_label_hr), dim=1)
    img_lr = resize_standin(img_lr, (slice_separation, 1), order=3)
    label_lr = resize_standin(patch_label_hr, (slice_separation, 1), order=0)
    img_lr = torch.cat((img_lr, label_lr), dim=1)
    img_hr = img_hr.permute(1, 2, 0, 3)
    img_lr = img_lr.permute(1, 2, 0, 3)
    if img_hr.shape[2] > 1 and rng.random() < 0.1:
        img_lr[:, 0:1] = torch.zeros_like(img_lr[:, 0:1])
    if img_hr.shape[2] > 1 and rng.random() < 0.1:
        img_lr[:, -1:] = torch.zeros_like(img_lr[:, -1:])
    if random_flip:
        for dim in (1, 2, 3):
            if rng.random() < 0.5:
                img_hr = img_hr.flip(dim)
                img_lr = img_lr.flip(dim)
    if rng.random() < 0.5:
        img_hr = img_hr.permute(0, 1, 3, 2).squeeze(3)
        img_lr = img_lr.permute(0, 1, 3, 2).squeeze(3)
    else:
        img_hr = img_hr.squeeze(2)
        img_lr = img_lr.squeeze(2)
    return img_lr, img_hr
