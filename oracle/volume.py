"""CPU restatement (numpy / torch-CPU) of the reference's driver-level functions -- TEST INFRASTRUCTURE.

  steps_for_sliding_window / sliding_window_slicers   <- utils/seg_utils.py:176-199, 229-238
  mirror_and_predict                                  <- utils/seg_utils.py:201-227
  sliding_window_logits                               <- utils/seg_utils.py:240-287 (fp16 accumulators, tile order)
  fba                                                 <- utils/fba.py:4-21
  rotate_vol_2d                                       <- utils/rotate.py:5-31
  get_pads / target_pad / crop                        <- utils/pad.py:5-32
  projected_size / ideal_size / find_integer_p / calc_slices_to_crop / get_patch   <- utils/patch_ops.py:6-64
  blur_same                                           <- F.conv2d(x, k, padding="same") at utils/train_set.py:325,332
  apply_to_vol_flavr                                  <- utils/sr_utils.py:102-135 (device-agnostic restatement)
  calculate_dice / evaluate_case_tensors              <- utils/seg_utils.py:730-734, 736-784 (everything after preprocess_image)
Pinned against the live reference functions (tests/test_oracle_vs_reference.py, run whenever /root/reference exists) and
the committed fixtures tests/golden/*.npz|json generated FROM the reference by oracle/make_golden.py.
"""
from __future__ import annotations

import itertools
import math

import numpy as np
import torch
import torch.nn.functional as F

from . import third_party as tp


def steps_for_sliding_window(image_size, tile_size, step):
    out = []
    for img, tile in zip(image_size, tile_size):
        n = int(np.ceil((img - tile) / (tile * step))) + 1
        delta = (img - tile) / (n - 1) if n > 1 else 99999999999
        out.append([int(np.round(delta * i)) for i in range(n)])
    return out


def sliding_window_slicers(image_size, patch_size, step=0.5):
    st = steps_for_sliding_window(image_size, patch_size, step)
    return [(slice(None), slice(a, a + patch_size[0]), slice(b, b + patch_size[1]), slice(c, c + patch_size[2]))
            for a in st[0] for b in st[1] for c in st[2]]


def mirror_and_predict(model, x, out_idx=None, deep_supervision=True):
    def pick(o):
        o = o[out_idx] if out_idx is not None else o
        return o[0] if (out_idx == 0 and deep_supervision) else o
    pred = pick(model(x))
    combos = [c for r in (1, 2, 3) for c in itertools.combinations((2, 3, 4), r)]
    for ax in combos:
        pred += torch.flip(pick(model(torch.flip(x, ax))), ax)
    pred /= len(combos) + 1
    return pred


def sliding_window_logits(data, slicers, network, out_idx=None, slice_sep=1, patch_size=(14, 320, 384), use_gaussian=False,
                          deep_supervision=True):
    logits = torch.zeros((2, data.shape[1] * slice_sep, data.shape[2], data.shape[3]), dtype=torch.half)
    npred = torch.zeros(logits.shape[1:], dtype=torch.half)
    g = tp.compute_gaussian(tuple(patch_size), sigma_scale=1. / 8, value_scaling_factor=10, device=torch.device("cpu")) \
        if use_gaussian else 1
    for sl in slicers:
        pred = mirror_and_predict(network, data[sl][None], out_idx, deep_supervision)[0]
        dst = (slice(None), slice(sl[1].start * slice_sep, sl[1].stop * slice_sep), sl[2], sl[3])
        logits[dst] += pred * g
        npred[dst[1:]] += g
    logits /= npred
    if torch.any(torch.isinf(logits)):
        raise RuntimeError("Encountered inf in predicted array")
    return logits


def fba(imgs, p="infinity"):
    spec = [np.fft.rfftn(v) for v in imgs]
    if p in ("infinity", "inf"):
        fused = np.max(spec, axis=0)          # numpy orders complex numbers lexicographically (real, then imag)
    else:
        mags = [np.abs(s) ** float(p) for s in spec]
        den = np.sum(mags, axis=0)
        fused = np.sum([m / den * s for m, s in zip(mags, spec)], axis=0)
    return np.fft.irfftn(fused).astype(np.float32)


def rotate_vol_2d(vol, angle):
    if angle in (0, 360):
        return vol
    if angle % 90 != 0 or abs(angle) > 270:
        raise NotImplementedError("Angles other than 90 degree rotations are not supported.")
    return torch.rot90(vol, k=int(angle // 90), dims=[0, 1])


def get_pads(target, d):
    if target <= d:
        return 0, 0
    lo = (target - d) // 2
    return lo, target - d - lo


def target_pad(img, target_dims, mode="reflect"):
    pads = tuple(get_pads(t, d) for t, d in zip(target_dims, img.shape))
    arr = img.numpy() if isinstance(img, torch.Tensor) else img
    out = np.pad(arr, pads, mode=mode)
    return (torch.Tensor(out) if isinstance(img, torch.Tensor) else out), pads


def crop(img, pads):
    return img[tuple(slice(lo or None, -hi if hi else None) for lo, hi in pads)]


def projected_size(n, p, s):
    return round((n + p) * (s / math.floor(s))) * math.floor(s) - round(p * s)


def ideal_size(n, s):
    return round(n * s)


def calc_slices_to_crop(p, s):
    return round(p * s)


def find_integer_p(n, s):
    p = 0
    while projected_size(n, p, s) != ideal_size(n, s) and p < 1000:
        p += 1
    return p if projected_size(n, p, s) == ideal_size(n, s) else 0


def get_patch_index(center, patch_size):
    starts = [c if p == 1 else c - p // 2 for c, p in zip(center, patch_size)]
    return tuple(slice(s, s + p) for s, p in zip(starts, patch_size))


def blur_same(x, kernel):
    return F.conv2d(x, kernel, padding="same")


def apply_to_vol_flavr(model, image, pred_out_idx=None):
    """image [Z, C, X, Y]: pad in-plane to a multiple of 16, sweep Z-1 windows of 4 slices (zero-padded at both ends),
    forward each (input cloned: the model mutates it), crop, concatenate along the slice axis -> [4(Z-1), C', X, Y]."""
    ox, oy = image.shape[2], image.shape[3]
    if ox % 16:
        image = torch.cat([image, image.new_zeros(image.shape[0], image.shape[1], 16 - ox % 16, image.shape[3])], 2)
    if oy % 16:
        image = torch.cat([image, image.new_zeros(image.shape[0], image.shape[1], image.shape[2], 16 - oy % 16)], 3)
    z = image.shape[0]
    outs = []
    for st in range(z - 1):
        if st == 0:
            win = image[0:3]
            win = torch.cat([win.new_zeros(4 - win.shape[0], *win.shape[1:]), win], 0)
        elif st == z - 2:
            win = image[st - 1:]
            win = torch.cat([win, win.new_zeros(4 - win.shape[0], *win.shape[1:])], 0)
        else:
            win = image[st - 1:st + 3]
        inp = win.permute(1, 0, 3, 2).unsqueeze(0).clone()
        with torch.no_grad():
            sr = model(inp)
            if pred_out_idx is not None and isinstance(sr, tuple):
                sr = sr[pred_out_idx]
        outs.append(sr.detach().cpu()[:, :, :, :oy, :ox])
    return torch.cat(outs, 2).squeeze(0).permute(1, 0, 2, 3)


def zscore_normalization(image):
    """utils/seg_utils.py:137-156 (tensor branch): per sample, channel 0 is normalised IN PLACE through a view (the caller's
    tensor changes), unbiased std floored at 1e-8; returns the stacked views [B, 1, ...]."""
    if isinstance(image, torch.Tensor):
        outs = []
        for i in range(image.shape[0]):
            v = image[i:i + 1, 0]
            mu, sd = v.mean(), v.std()
            v -= mu
            v /= max(sd, 1e-8)
            outs.append(v)
        return torch.stack(outs, 0)
    image = image.astype(np.float32, copy=False)
    mu, sd = image.mean(), image.std()
    image -= mu
    image /= max(sd, 1e-8)
    return image


def sr_volume_orientations(model, image, angles=(0,), pred_out_idx=0):
    """utils/sr_utils.py:157-173: rotate, permute to (hr, C, lr, hr), window sweep, permute back, un-rotate, mean."""
    preds = []
    for angle in angles:
        rot = rotate_vol_2d(image, angle).permute(0, 3, 2, 1)
        res = apply_to_vol_flavr(model, rot, pred_out_idx).permute(0, 3, 1, 2)
        preds.append(rotate_vol_2d(res, -angle))
    return torch.mean(torch.stack(preds), dim=0)


def calculate_dice(prediction, ground_truth, smooth=1e-5):
    """utils/seg_utils.py:730-734."""
    prediction = np.asarray(prediction).flatten()
    ground_truth = np.asarray(ground_truth).flatten()
    intersection = np.sum(prediction * ground_truth)
    return (2. * intersection + smooth) / (np.sum(prediction) + np.sum(ground_truth) + smooth)


def evaluate_case_tensors(model, lr_data, lr_label, slice_separation, patch_size, get_HR_results=False):
    """utils/seg_utils.py:736-784 from the padding on (preprocess_image is NIfTI IO): pad, sliding window over output 0 with the
    Gaussian, crop, softmax, argmax, Dice; optional HR branch over output 1 without the Gaussian (the reference's defaults)."""
    model.eval()
    lr_data, revert = tp.pad_nd_image(lr_data, patch_size, 'constant', {'value': 0}, True, None)
    with torch.no_grad():
        slicers = sliding_window_slicers(lr_data.shape[1:], list(patch_size))
        logits = sliding_window_logits(lr_data, slicers, model, 0, 1, patch_size, use_gaussian=True, deep_supervision=False)
    prediction = logits[tuple([slice(None), *revert[1:]])].squeeze(0)
    prob = torch.softmax(prediction.float(), dim=0).numpy()
    prediction_lr = prob.argmax(0).astype('uint8')
    dice_lr = calculate_dice(prediction_lr, lr_label.squeeze(0).numpy().astype('uint8'))
    if get_HR_results:
        with torch.no_grad():
            hr = sliding_window_logits(lr_data, slicers, model, 1, slice_separation,
                                       [patch_size[0] * slice_separation, patch_size[1], patch_size[2]])
        prediction_hr = torch.argmax(hr, dim=0).squeeze(0).numpy().astype('uint8')
    else:
        prediction_hr = prediction_lr
    return prediction_lr, prediction_hr, dice_lr
