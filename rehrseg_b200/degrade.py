"""GPU-side sample synthesis of the SR stage.

The reference builds every training sample of the self-SR stage on the CPU in `TrainSetMultiple.__getitem__`
(utils/train_set.py:337-434) with `num_workers=0` (train_all.py:371-378): random crop of the HR volume and of its pre-blurred copy,
padding, cubic resampling to the low resolution (`resize(img, (slice_separation, 1), order=3)`, nearest for the label), slice
dropout, flips, final permutation.  At B200 speeds (> 300 samples/s for the network) that loader is the
bottleneck.  Here the volumes stay resident in HBM and the same statements run as device ops; the random decisions are drawn from
Python's `random` in the reference's order, so a seeded run reproduces the reference's sample stream.

* `blur_prefilter`  = `load_img`'s two `F.conv2d(..., padding="same")` calls (:321-333) on `volume_ops.blur_along_x` (rehr_blur1d)
* `resize`          = the stand-in for the third-party `resize.pytorch.resize` defined in oracle/degrade.py (its source is not
                      available: "parity unpinned" for this one function) on `rehr_resample_axis`
* `SRTrainSampler`  = the dataset object: `sample(i)` mirrors `__getitem__`, `batch(n)` is what the DataLoader's collate returns

Parity: tests/test_degrade_gpu.py against tests/golden/sr_degrade.npz, which oracle/make_golden.py produced by running the
reference's OWN `__getitem__` with the stand-in injected."""
from __future__ import annotations

import random as _random
from typing import List, Optional, Sequence, Tuple

import torch

from . import functional as F_
from ._lib import RehrError, check, lib, ptr, stream_ptr
from .volume_ops import blur_along_x, get_pads


def _resample_axis(x: torch.Tensor, axis: int, d: float, order: int) -> torch.Tensor:
    if not x.is_cuda:
        raise RehrError("rehrseg_b200 ops need CUDA tensors (no CPU fallback)")
    xs = x.float().contiguous()
    n_in = xs.shape[axis]
    n_out = int(round(n_in / d))
    outer = 1
    for v in xs.shape[:axis]:
        outer *= v
    inner = 1
    for v in xs.shape[axis + 1:]:
        inner *= v
    y = torch.empty((*xs.shape[:axis], n_out, *xs.shape[axis + 1:]), dtype=torch.float32, device=xs.device)
    if y.numel():
        check(lib().rehr_resample_axis(ptr(xs), ptr(y), outer, n_in, n_out, inner, float(d), int(order), stream_ptr()), "resample_axis")
        F_._count()
    return y


def resize(x: torch.Tensor, dxy: Sequence[float], order: int = 3) -> torch.Tensor:
    """`resize.pytorch.resize(image[B, C, X, Y], (dx, dy), order)` as called at utils/train_set.py:395-396,516 -- the stand-in
    definition of oracle/degrade.py: n_out = round(n / d), sample i at (i + 0.5) d - 0.5, cubic convolution A = -0.75 (order 3) or
    nearest (order 0), clamped borders.  fp32."""
    if x.dim() != 4 or len(dxy) != 2:
        raise RehrError("resize expects a [B, C, X, Y] tensor and two step factors")
    if order not in (0, 3):
        raise RehrError("resize: order 3 (cubic) and 0 (nearest) are implemented")
    out = x
    for axis, d in ((2, float(dxy[0])), (3, float(dxy[1]))):
        if d != 1.0:
            out = _resample_axis(out, axis, d, order)
    return out.float() if out is x else out


def blur_prefilter(image_xyzc: torch.Tensor, kernel: torch.Tensor) -> Tuple[torch.Tensor, torch.Tensor]:
    """`load_img`, utils/train_set.py:321-333: channel 0 of `image` [X, Y, Z, C] blurred along X as [Z,1,X,Y] and along Y as [Z,1,Y,X]
    (`F.conv2d(x, kernel[1,1,L,1], padding="same")` on both)."""
    image_x = image_xyzc.permute(2, 3, 0, 1)[:, 0:1].contiguous()
    image_y = image_xyzc.permute(2, 3, 1, 0)[:, 0:1].contiguous()
    return blur_along_x(image_x, kernel), blur_along_x(image_y, kernel)


def _target_pad_const(img: torch.Tensor, target_dims) -> torch.Tensor:
    pads = tuple(get_pads(t, d) for t, d in zip(target_dims, img.shape))     # utils/pad.py:14-20, mode="constant"
    flat = []
    for b, a in reversed(pads):
        flat += [b, a]
    return torch.nn.functional.pad(img, flat) if any(flat) else img


class SRTrainSampler:
    """`TrainSetMultiple` with the volumes resident on the GPU (`train_transform=None`, the batchgenerators branch is out of scope).
    `add_subject` = one iteration of the preload loop (utils/train_set.py:284-298); `sample` = `__getitem__`."""

    def __init__(self, patch_size: Sequence[int], slice_separation: float, blur: bool = True, random_flip: bool = False,
                 blur_kernel: Optional[torch.Tensor] = None, device=None):
        self.patch_size = list(patch_size)
        self.slice_separation = float(slice_separation)
        self.blur, self.random_flip = bool(blur), bool(random_flip)
        self.blur_kernel = blur_kernel
        self.device = torch.device(device) if device is not None else torch.device("cuda", torch.cuda.current_device())
        self.imgs_hr: List[torch.Tensor] = []
        self.labels_hr: List[torch.Tensor] = []
        self.imgs_filtered_x: List[Optional[torch.Tensor]] = []
        self.imgs_filtered_y: List[Optional[torch.Tensor]] = []

    def __len__(self) -> int:
        return len(self.imgs_hr)

    def add_subject(self, img_hr, label_hr, filtered_x=None, filtered_y=None) -> None:
        """img_hr [X,Y,Z,1] float, label_hr [X,Y,Z,1]; the blurred copies are computed here (rehr_blur1d) unless given."""
        img = torch.as_tensor(img_hr).to(self.device, torch.float32)
        lab = torch.as_tensor(label_hr).to(self.device)
        if self.blur and (filtered_x is None or filtered_y is None):
            if self.blur_kernel is None:
                raise RehrError("SRTrainSampler(blur=True) needs blur_kernel or pre-filtered volumes")
            fx, fy = blur_prefilter(torch.cat((img, lab.to(torch.float32)), dim=-1), self.blur_kernel)
        else:
            fx = None if filtered_x is None else torch.as_tensor(filtered_x).to(self.device, torch.float32)
            fy = None if filtered_y is None else torch.as_tensor(filtered_y).to(self.device, torch.float32)
        self.imgs_hr.append(img)
        self.labels_hr.append(lab)
        self.imgs_filtered_x.append(fx)
        self.imgs_filtered_y.append(fy)

    def sample(self, i: int, rng=_random) -> Tuple[torch.Tensor, torch.Tensor]:
        """(img_lr, img_hr) of utils/train_set.py:337-434; `rng` is consumed exactly like the reference's `random`."""
        img_hr, label_hr = self.imgs_hr[i], self.labels_hr[i]
        img_lr = None
        if self.blur:
            if rng.random() < 0.5:
                img_hr, label_hr = img_hr.permute(1, 0, 2, 3), label_hr.permute(1, 0, 2, 3)
                img_lr = self.imgs_filtered_y[i]
            else:
                img_lr = self.imgs_filtered_x[i]
        elif rng.random() < 0.5:
            img_hr, label_hr = img_hr.permute(1, 0, 2, 3), label_hr.permute(1, 0, 2, 3)
        ps = self.patch_size
        x_0 = rng.randint(0, max(img_hr.shape[0] - ps[0], 0))
        y_0 = rng.randint(0, max(img_hr.shape[1] - ps[1], 0))
        z_0 = rng.randint(0, max(img_hr.shape[2] - ps[2], 0))
        img_hr = img_hr[x_0:x_0 + ps[0], y_0:y_0 + ps[1], z_0:z_0 + ps[2], :]
        patch_label_hr = label_hr[x_0:x_0 + ps[0], y_0:y_0 + ps[1], z_0:z_0 + ps[2], :].to(torch.float32)
        img_hr = img_hr.permute(2, 3, 0, 1)                       # z, channel, x, y
        patch_label_hr = patch_label_hr.permute(2, 3, 0, 1)
        target_shape = [max(s, p) for s, p in zip(img_hr.shape, (ps[2], 1, ps[0], ps[0]))]   # ps[0] twice: the reference's line 362
        img_hr = _target_pad_const(img_hr, target_shape)
        patch_label_hr = _target_pad_const(patch_label_hr, target_shape)
        if self.blur:
            img_lr = img_lr[z_0:z_0 + ps[2], :, x_0:x_0 + ps[0], y_0:y_0 + ps[1]]
            img_lr = _target_pad_const(img_lr, target_shape)
        else:
            img_lr = img_hr
        img_hr = torch.cat((img_hr, patch_label_hr), dim=1)
        img_lr = resize(img_lr, (self.slice_separation, 1), order=3)          # simulate the LR image
        label_lr = resize(patch_label_hr, (self.slice_separation, 1), order=0)
        img_lr = torch.cat((img_lr, label_lr), dim=1)
        img_hr = img_hr.permute(1, 2, 0, 3)                       # channel, x, z, y
        img_lr = img_lr.permute(1, 2, 0, 3)
        if img_hr.shape[2] > 1 and rng.random() < 0.1:
            img_lr[:, 0:1] = 0
        if img_hr.shape[2] > 1 and rng.random() < 0.1:
            img_lr[:, -1:] = 0
        if self.random_flip:
            for dim in (1, 2, 3):
                if rng.random() < 0.5:
                    img_hr, img_lr = img_hr.flip(dim), img_lr.flip(dim)
        if rng.random() < 0.5:
            img_hr = img_hr.permute(0, 1, 3, 2).squeeze(3)
            img_lr = img_lr.permute(0, 1, 3, 2).squeeze(3)
        else:
            img_hr, img_lr = img_hr.squeeze(2), img_lr.squeeze(2)
        return img_lr, img_hr

    def batch(self, indices: Sequence[int], rng=_random) -> Tuple[torch.Tensor, torch.Tensor]:
        """What `DataLoader(dataset, batch_size=len(indices))` yields (default collate = stack), on the device."""
        pairs = [self.sample(i, rng) for i in indices]
        return torch.stack([p[0] for p in pairs]).contiguous(), torch.stack([p[1] for p in pairs]).contiguous()
