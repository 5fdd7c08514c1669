"""Volume-level helpers of the self-SR pipeline -- drop-ins for utils/rotate.py, utils/pad.py, utils/patch_ops.py,
utils/fba.py, the blur application `F.conv2d(x, k[1,1,L,1], padding="same")` (utils/train_set.py:325,332;
utils/sr_utils.py:272,276,302) and the orientation mean (utils/sr_utils.py:65,173,217) of the reference.

Index math (pads, crops, slice-padding search, patch slicing) is plain Python/numpy, bit-exact with the reference
(including Python's round-half-to-even).  The tensor work runs in the HBM-bound kernels of librehrseg_b200.so
(rehr_rot90, rehr_blur1d, rehr_fba_combine, rehr_mean_stack); FFTs go through cuFFT (torch.fft), a library call the
spec allows for plain transforms.
"""
from __future__ import annotations

import ctypes as C
from math import floor
from typing import List, Sequence

import numpy as np
import torch

from ._lib import RehrError, check, lib, ptr, stream_ptr
from . import functional as F_


# ----------------------------------------------------------------------------------------------------------------
# utils/rotate.py
# ----------------------------------------------------------------------------------------------------------------
_ANGLE_TO_K = {90: 1, -90: -1, 180: 2, -180: -2, 270: 3, -270: -3}


def rotate_vol_2d(vol: torch.Tensor, angle):
    """utils/rotate.py:5-31: rot90 over dims (0, 1) for multiples of 90 degrees, identity (same object) for 0 / 360,
    NotImplementedError otherwise.  A pure permutation: bit-exact."""
    if angle == 0 or angle == 360:
        return vol
    if angle not in _ANGLE_TO_K:
        raise NotImplementedError("Angles other than 90 degree rotations are not supported.")
    if not vol.is_cuda:
        raise RehrError("rehrseg_b200 ops need CUDA tensors (no CPU fallback)")
    k = _ANGLE_TO_K[angle] % 4
    src = vol.contiguous()
    X, Y = src.shape[0], src.shape[1]
    inner = 1
    for s in src.shape[2:]:
        inner *= s
    out_shape = ((Y, X) if k % 2 else (X, Y)) + tuple(src.shape[2:])
    dst = torch.empty(out_shape, dtype=src.dtype, device=src.device)
    if src.numel():
        check(lib().rehr_rot90(ptr(src), ptr(dst), X, Y, inner * src.element_size(), k, stream_ptr()), "rot90")
        F_._count()
    return dst


# ----------------------------------------------------------------------------------------------------------------
# utils/pad.py
# ----------------------------------------------------------------------------------------------------------------
def get_pads(target_dim, d):
    """utils/pad.py:5-11."""
    if target_dim <= d:
        return 0, 0
    p = (target_dim - d) // 2
    return p, target_dim - d - p


def target_pad(img, target_dims, mode="reflect"):
    """utils/pad.py:14-20: centre-pad `img` up to target_dims; returns (padded, pads).  CUDA tensors are padded on the
    device (the reference round-trips through numpy); reflect / constant / edge modes."""
    pads = tuple(get_pads(t, d) for t, d in zip(target_dims, img.shape))
    if isinstance(img, torch.Tensor):
        if not img.is_cuda:
            return torch.Tensor(np.pad(img.numpy(), pads, mode=mode)), pads
        tmode = {"reflect": "reflect", "constant": "constant", "edge": "replicate"}.get(mode)
        if tmode is None:
            raise RehrError(f"target_pad: mode {mode!r} is not implemented on the device")
        out = img
        # torch pads at most the 3 trailing dims of a batched tensor per call for non-constant modes: do one axis at a time
        for ax, (b, a) in enumerate(pads):
            if b == 0 and a == 0:
                continue
            moved = out.movedim(ax, -1)
            shp = moved.shape
            flat = moved.reshape(1, -1, shp[-1])
            flat = torch.nn.functional.pad(flat, (b, a), mode=tmode)
            out = flat.reshape(*shp[:-1], shp[-1] + a + b).movedim(-1, ax)
        return out.contiguous(), pads
    return np.pad(img, pads, mode=mode), pads


def format_pads(pads):
    """utils/pad.py:23-27."""
    st = pads[0] if pads[0] != 0 else None
    en = -pads[1] if pads[1] != 0 else None
    return slice(st, en)


def crop(img, pads):
    """utils/pad.py:30-32."""
    return img[tuple(map(format_pads, pads))]


# ----------------------------------------------------------------------------------------------------------------
# utils/patch_ops.py (integer helpers)
# ----------------------------------------------------------------------------------------------------------------
def projected_size(n_slices, p, scale):
    """utils/patch_ops.py:6-12 (Python round: half to even)."""
    scale_tilde = scale / floor(scale)
    return round((n_slices + p) * scale_tilde) * floor(scale) - round(p * scale)


def calc_slices_to_crop(p, scale):
    """utils/patch_ops.py:15-16."""
    return round(p * scale)


def ideal_size(n_slices, scale):
    """utils/patch_ops.py:19-24."""
    return round(n_slices * scale)


def find_integer_p(n_slices, s):
    """utils/patch_ops.py:27-46: smallest p <= 1000 whose padded-then-cropped slice count equals round(n*s); 0 if none."""
    want = ideal_size(n_slices, s)
    for p in range(0, 1001):
        if projected_size(n_slices, p, s) == want:
            return p
    return 0


def get_patch(img_rot, patch_center, patch_size, return_idx=False):
    """utils/patch_ops.py:49-64: start = c - p // 2 (c itself when p == 1)."""
    sts = [c - p // 2 if p != 1 else c for c, p in zip(patch_center, patch_size)]
    idx = tuple(slice(st, st + p) for st, p in zip(sts, patch_size))
    if return_idx:
        return idx
    return img_rot[idx].squeeze()


# ----------------------------------------------------------------------------------------------------------------
# blur degradation
# ----------------------------------------------------------------------------------------------------------------
def blur_along_x(x: torch.Tensor, kernel: torch.Tensor) -> torch.Tensor:
    """F.conv2d(x[Z,1,X,Y], kernel[1,1,L,1], padding="same"): L-tap cross-correlation along dim 2, zero padded, fp32.
    For even L PyTorch pads (L-1)//2 on the left and the rest on the right; the kernel does the same."""
    if x.dim() != 4 or x.shape[1] != 1 or kernel.dim() != 4 or kernel.shape[:2] != (1, 1) or kernel.shape[3] != 1:
        raise RehrError("blur_along_x expects x [Z,1,X,Y] and kernel [1,1,L,1] (utils/blur_kernel_ops.py:7-18)")
    if not x.is_cuda:
        raise RehrError("rehrseg_b200 ops need CUDA tensors (no CPU fallback)")
    xs = x.contiguous().float()
    taps = kernel.reshape(-1).to(device=x.device, dtype=torch.float32).contiguous()
    L = taps.numel()
    if L > 65:
        raise RehrError("blur kernels longer than 65 taps are not implemented")
    y = torch.empty_like(xs)
    Z, _, X, Y = xs.shape
    if xs.numel():
        check(lib().rehr_blur1d(ptr(xs), ptr(taps), L, ptr(y), Z, X, Y, stream_ptr()), "blur1d")
        F_._count()
    return y


# ----------------------------------------------------------------------------------------------------------------
# orientation fusion
# ----------------------------------------------------------------------------------------------------------------
def _ptr_array(tensors: Sequence[torch.Tensor]):
    arr = (C.c_void_p * len(tensors))(*[t.data_ptr() for t in tensors])
    return arr


def mean_fuse(vols: Sequence[torch.Tensor]) -> torch.Tensor:
    """torch.mean(torch.stack(vols), dim=0) (utils/sr_utils.py:65,173,217) without materialising the stack."""
    if not vols:
        raise RehrError("mean_fuse: empty list")
    vs = [v.contiguous().float() for v in vols]
    if any(v.shape != vs[0].shape or not v.is_cuda for v in vs):
        raise RehrError("mean_fuse: volumes must be CUDA tensors of one shape")
    if len(vs) > 16:
        raise RehrError("mean_fuse: at most 16 volumes")
    out = torch.empty_like(vs[0])
    check(lib().rehr_mean_stack(_ptr_array(vs), len(vs), ptr(out), out.numel(), stream_ptr()), "mean_stack")
    F_._count()
    return out


def fba(imgs: List, p="infinity"):
    """utils/fba.py:4-21: Fourier Burst Accumulation.  rfftn of every volume, combined per frequency bin -- p = "infinity":
    numpy's `max` on complex numbers, i.e. the LEXICOGRAPHIC (real, then imaginary) maximum, not the largest magnitude;
    finite p: sum_i w_i v_i with w_i = |v_i|^p / sum_j |v_j|^p -- then irfftn WITHOUT an explicit size (so an odd last
    dimension comes back one shorter, as in the reference) and a float32 result.  Accepts numpy arrays (returns numpy,
    like the reference) or CUDA tensors (returns a CUDA tensor)."""
    if len(imgs) == 0:
        raise ValueError("need at least one array to stack")  # np.max([]) in the reference raises ValueError too
    as_numpy = not isinstance(imgs[0], torch.Tensor)
    dev = torch.device("cuda", torch.cuda.current_device())
    vols = [torch.as_tensor(np.ascontiguousarray(v), dtype=torch.float32).to(dev) if as_numpy else v.to(dev, torch.float32)
            for v in imgs]
    if len(vols) > 16:
        raise RehrError("fba: at most 16 volumes")
    specs = [torch.fft.rfftn(v).contiguous() for v in vols]  # cuFFT R2C, complex64 (numpy >= 2 also computes in complex64)
    out = torch.empty_like(specs[0])
    if p == "infinity" or p == "inf":
        pv = -1.0
    else:
        pv = float(p)
        if pv < 0:
            raise RehrError("fba: p must be >= 0 or 'infinity'")
    views = [torch.view_as_real(s) for s in specs]
    check(lib().rehr_fba_combine(_ptr_array(views), len(views), pv, ptr(torch.view_as_real(out)), out.numel(), stream_ptr()),
          "fba_combine")
    F_._count()
    res = torch.fft.irfftn(out).to(torch.float32)
    return res.cpu().numpy() if as_numpy else res


def get_random_centers(imgs_rot, patch_size, n_patches, weighted=True):
    """utils/patch_ops.py:67-113: `n_patches` (rotation index, centre) pairs.  Each patch first draws its rotation uniformly; per
    rotation the centre coordinates are drawn axis by axis from the marginals of a gradient-magnitude map (sum over axes of
    sqrt|d/dx_k| of the sigma-1 Gaussian-smoothed image, zeroed within p//2 + 1 voxels of the borders of every axis with p > 1) or
    uniformly when `weighted` is False; the list is shuffled at the end.  Host-side numpy / scipy like the reference (index
    sampling for the CPU dataset); it consumes numpy's GLOBAL random stream in the reference's order (randint, one choice per axis
    and rotation, shuffle), so the same `np.random.seed` gives the same centres."""
    import numpy as np
    from scipy.ndimage import gaussian_filter
    which = np.random.randint(0, len(imgs_rot), size=n_patches)
    centers = []
    for i, img in enumerate(imgs_rot):
        count = int(np.sum(which == i))
        if weighted:
            mag = np.sum([np.sqrt(np.abs(d)) for d in np.gradient(gaussian_filter(img, 1.0))], axis=0)
            for axis, p in enumerate(patch_size):
                if axis < mag.ndim and p > 1:           # no patch centre may sit where the patch would leave the image
                    view = np.swapaxes(mag, 0, axis)
                    view[: p // 2 + 1] = 0.0
                    view[-p // 2 - 1:] = 0.0
            joint = mag / mag.sum()
            probs = []
            for axis in range(joint.ndim):
                m = joint.sum(axis=tuple(k for k in range(joint.ndim) if k != axis))
                probs.append(m / m.sum())
        else:
            probs = [None] * img.ndim
        draws = [np.random.choice(np.arange(0, dim), size=count, p=probs[axis]) for axis, dim in enumerate(img.shape)]
        centers.extend((i, tuple(c)) for c in zip(*draws))
    np.random.shuffle(centers)
    return centers


# ----------------------------------------------------------------------------------------------------------------
# post-processing of the SR stage outputs (the tensor parts of utils/sr_utils.py:244-304, after parse_image's file IO)
# ----------------------------------------------------------------------------------------------------------------
def zeroonenorm(data: torch.Tensor) -> torch.Tensor:
    """utils/sr_utils.py:279-282: min-max normalise to [0, 255] (fp32)."""
    data = data.float()
    lo, hi = data.amin(), data.amax()
    return (data - lo) / (hi - lo) * 255.0


def postprocess_flavr_volume(image: torch.Tensor, blur_kernel: torch.Tensor) -> torch.Tensor:
    """`postprocess_flavr` (utils/sr_utils.py:284-304) between the two parse_image reads and the return: min-max normalise the SR
    volume [X, Y, Z] to [0, 255], move z first ([Z, 1, X, Y]), blur along x with the slice-profile kernel [1, 1, L, 1]
    (F.conv2d(..., padding="same") -> rehr_blur1d) and move z back.  Returns [X, Y, Z] fp32 on the GPU."""
    if image.dim() != 3:
        raise RehrError("postprocess_flavr_volume expects the [X, Y, Z] SR volume")
    vol = zeroonenorm(image.cuda() if not image.is_cuda else image)
    z_first = vol.permute(2, 0, 1).unsqueeze(1)                        # z, 1, x, y
    return blur_along_x(z_first, blur_kernel).squeeze(1).permute(1, 2, 0)


def postprocess_smore_volume(image: torch.Tensor, blur_kernel: torch.Tensor):
    """`postprocess_smore` (utils/sr_utils.py:244-277) after the volume [X, Y, Z, 2] (image, label) has been assembled: returns
    (img_hr [X,Y,Z,1], label_hr uint8 [X,Y,Z,1], image_x_rgb [Z,1,X,Y], image_y_rgb [Z,1,Y,X]) -- the two in-plane orientations of
    the image channel blurred along their first in-plane axis."""
    if image.dim() != 4 or image.shape[-1] < 2:
        raise RehrError("postprocess_smore_volume expects [X, Y, Z, 2] (image, label)")
    image = image.cuda() if not image.is_cuda else image
    img_hr = image[..., :1]
    label_hr = image[..., 1:].to(torch.uint8)
    image_x = image.permute(2, 3, 0, 1)[:, 0:1]                        # z, channel, x, y
    image_y = image.permute(2, 3, 1, 0)[:, 0:1]                        # z, channel, y, x
    return img_hr, label_hr, blur_along_x(image_x, blur_kernel), blur_along_x(image_y, blur_kernel)
