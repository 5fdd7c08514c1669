"""Sliding-window Gaussian-blended inference -- drop-in for utils/seg_utils.py:176-287 (+ the driver part of
`evaluate_case`, utils/seg_utils.py:741-763) of the reference.

Same function names, argument meaning, tile order (x-major, z-minor), accumulator dtype (fp16) and error behaviour
(`RuntimeError` on inf logits, the reference's `assert`s) as the reference; the per-tile network forwards run on the
sm_100a engine and the blend `logits[sl] += pred * g; n[sl] += g; logits /= n` runs in the `rehr_sw_*` kernels, which
reproduce ATen's half arithmetic bit for bit.  Host-side integer logic (tile origins, padding) is plain Python and is
exercised on CPU by tests/test_host_logic.py.

Multi-GPU: `predict_sliding_window_sharded` deals the tiles round-robin over the ranks of a
torch.distributed group; each rank blends its tiles into its own buffers and one all-reduce merges them.
"""
from __future__ import annotations

import ctypes as C
import itertools
import os
from functools import lru_cache
from typing import Callable, List, Optional, Sequence, Tuple

import numpy as np
import torch

from ._lib import RehrError, check, lib, ptr, stream_ptr
from . import functional as F_


# ----------------------------------------------------------------------------------------------------------------
# integer tile geometry (bit-exact with the reference)
# ----------------------------------------------------------------------------------------------------------------
def compute_steps_for_sliding_window(image_size, tile_size, tile_step_size) -> List[List[int]]:
    """utils/seg_utils.py:176-199.  n = ceil((img - tile) / (tile * step)) + 1 origins per axis, spread evenly:
    round(i * (img - tile) / (n - 1)) with numpy's round-half-to-even."""
    # the reference asserts on a (always truthy) list here; the real precondition is enforced instead
    if any(i < j for i, j in zip(image_size, tile_size)):
        raise AssertionError("image size must be as large or larger than patch_size")
    assert 0 < tile_step_size <= 1, 'step_size must be larger than 0 and smaller or equal to 1'
    steps = []
    for img, tile in zip(image_size, tile_size):
        target = tile * tile_step_size
        n = int(np.ceil((img - tile) / target)) + 1
        span = img - tile
        actual = span / (n - 1) if n > 1 else 99999999999
        steps.append([int(np.round(actual * i)) for i in range(n)])
    return steps


def _internal_get_sliding_window_slicers(image_size, patch_size=[14, 320, 384], tile_step_size=0.5):
    """utils/seg_utils.py:229-238: slicers (slice(None), x, y, z) in x-major / z-minor order."""
    steps = compute_steps_for_sliding_window(image_size, patch_size, tile_step_size)
    slicers = []
    for sx in steps[0]:
        for sy in steps[1]:
            for sz in steps[2]:
                slicers.append(tuple([slice(None), *[slice(si, si + ti) for si, ti in zip((sx, sy, sz), patch_size)]]))
    return slicers


@lru_cache(maxsize=2)
def compute_gaussian(tile_size: Tuple[int, ...], sigma_scale: float = 1. / 8, value_scaling_factor: float = 1,
                     dtype=torch.float16, device=torch.device("cuda", 0)) -> torch.Tensor:
    """nnunetv2 2.3.1 `compute_gaussian` (imported by the reference at utils/seg_utils.py:11, used at :261-263): a delta at
    the tile centre filtered by scipy's zero-padded Gaussian (sigma = tile * sigma_scale), scaled to a maximum of
    `value_scaling_factor`, cast to `dtype`, zeros replaced by the smallest non-zero entry.  lru_cached like the original
    (the reference calls `.cache_clear()`, utils/seg_utils.py:285)."""
    from scipy.ndimage import gaussian_filter
    tmp = np.zeros(tile_size)
    tmp[tuple(i // 2 for i in tile_size)] = 1
    g = torch.from_numpy(gaussian_filter(tmp, [i * sigma_scale for i in tile_size], 0, mode="constant", cval=0))
    g = g / torch.max(g) * value_scaling_factor
    g = g.to(device=device, dtype=dtype)
    g[g == 0] = torch.min(g[g != 0])
    return g


# The reference clears nnU-Net's lru_cache after every volume (utils/seg_utils.py:285), which re-runs scipy's Gaussian filter on
# the host for every case (~0.9 s for a 128^3 tile against ~0.5 s of GPU work for a whole 256^3 volume).  The importance map only
# depends on (tile, sigma, scale, dtype, device), so the drivers below keep it in their own memo; `compute_gaussian` and its
# `.cache_clear()` stay what the reference expects.
_gaussian_memo: dict = {}


def importance_map(tile_size, sigma_scale: float = 1. / 8, value_scaling_factor: float = 10, dtype=torch.float16, device=None):
    key = (tuple(int(t) for t in tile_size), float(sigma_scale), float(value_scaling_factor), dtype, str(device))
    g = _gaussian_memo.get(key)
    if g is None:
        g = compute_gaussian(key[0], sigma_scale=sigma_scale, value_scaling_factor=value_scaling_factor, dtype=dtype, device=device)
        if len(_gaussian_memo) >= 8:
            _gaussian_memo.clear()
        _gaussian_memo[key] = g
    return g


def pad_nd_image(image, new_shape, mode="constant", kwargs=None, return_slicer=False):
    """acvl_utils 0.2 `pad_nd_image` as the reference calls it (utils/seg_utils.py:741): pad the trailing len(new_shape)
    dims up to new_shape, below = diff // 2, above = diff // 2 + diff % 2; returns (padded, slicer)."""
    kwargs = kwargs or {}
    old = list(image.shape)
    nd = len(new_shape)
    tgt = old[:-nd] + [max(o, n) for o, n in zip(old[-nd:], new_shape)]
    pads = [[(t - o) // 2, (t - o) // 2 + (t - o) % 2] for t, o in zip(tgt, old)]
    if any(b or a for b, a in pads):
        if isinstance(image, torch.Tensor):
            flat = [v for pr in pads[::-1] for v in pr]
            res = torch.nn.functional.pad(image, flat, mode=mode, **kwargs)
        else:
            np_kwargs = {"constant_values": kwargs["value"]} if (mode == "constant" and "value" in kwargs) else {}
            res = np.pad(image, pads, mode, **np_kwargs)
    else:
        res = image
    if not return_slicer:
        return res
    return res, tuple(slice(p[0], res.shape[i] - p[1]) for i, p in enumerate(pads))


# ----------------------------------------------------------------------------------------------------------------
# per-tile prediction with mirror test-time augmentation
# ----------------------------------------------------------------------------------------------------------------
MIRROR_AXES = (0, 1, 2)  # hard-coded in the reference (utils/seg_utils.py:202)
SW_PAIR_VARIANTS = os.environ.get("REHR_SW_PAIRS", "1") != "0"   # graph-replayed tile forwards on several mirror variants at a time
SW_SHARD_GROUP = int(os.environ.get("REHR_SW_SHARD_GROUP", "4"))   # sharded driver: this rank's (tile, variant) units per forward
SW_VARIANT_GROUP = int(os.environ.get("REHR_SW_GROUP", "8"))      # ... this many (1, 2, 4 or 8; measured on C3: 0.512 / 0.446 / 0.432 / 0.429 s per volume)


def mirror_axes_combinations(ndim: int = 5):
    return [c for i in range(len(MIRROR_AXES)) for c in itertools.combinations([m + 2 for m in MIRROR_AXES], i + 1)]


def _select(out, out_idx, deep_supervision):
    p = out[out_idx] if out_idx is not None else out
    if out_idx == 0 and deep_supervision:
        p = p[0]
    return p


def _internal_maybe_mirror_and_predict(model, x, out_idx=None, deep_supervision=True, save=False,
                                       accum_dtype: Optional[torch.dtype] = None, pair_model=None, group: int = 2) -> torch.Tensor:
    """utils/seg_utils.py:201-227: model(x) plus the 7 flipped variants, averaged.  The reference runs this under fp16
    autocast (utils/seg_utils.py:743-744), so the running sum is fp16 there; `accum_dtype` selects that (default: the
    dtype the model returns).  `pair_model` (a callable on a batch of `group` tiles): the 8 variants run `group` per forward
    (InstanceNorm is per sample, so a variant's logits do not depend on its batch mates); they are summed in the reference's order."""
    assert max(MIRROR_AXES) <= x.ndim - 3, 'mirror_axes does not match the dimension of the input!'
    combos = mirror_axes_combinations()
    if pair_model is not None and x.shape[0] == 1 and (len(combos) + 1) % group == 0:
        variants = [()] + [tuple(c) for c in combos]
        prediction = None
        for k in range(0, len(variants), group):
            xs = [torch.flip(x, v) if v else x for v in variants[k:k + group]]
            pg = _select(pair_model(torch.cat(xs, dim=0)), out_idx, deep_supervision)
            for j, axes in enumerate(variants[k:k + group]):
                p = pg[j:j + 1]
                if prediction is None:
                    prediction = p.to(accum_dtype) if accum_dtype is not None else p.clone()
                else:
                    prediction += torch.flip(p, axes).to(prediction.dtype)
        prediction /= len(variants)
        return prediction
    prediction = _select(model(x), out_idx, deep_supervision)
    prediction = prediction.to(accum_dtype) if accum_dtype is not None else prediction.clone()
    for axes in combos:
        p = _select(model(torch.flip(x, (*axes,))), out_idx, deep_supervision)
        prediction += torch.flip(p, (*axes,)).to(prediction.dtype)
    prediction /= (len(combos) + 1)
    return prediction


# ----------------------------------------------------------------------------------------------------------------
# blend
# ----------------------------------------------------------------------------------------------------------------
def sw_accumulate(predicted_logits: torch.Tensor, n_predictions: Optional[torch.Tensor], prediction: torch.Tensor, gaussian,
                  origin: Sequence[int]) -> None:
    """logits[:, o:o+t] += prediction * gaussian ; n[o:o+t] += gaussian  (utils/seg_utils.py:275-276), fp16 accumulators.
    `n_predictions=None` blends the logits only."""
    if predicted_logits.dtype != torch.float16 or (n_predictions is not None and n_predictions.dtype != torch.float16):
        raise RehrError("sliding-window accumulators are fp16 as in the reference (utils/seg_utils.py:256-259)")
    cch, vd, vh, vw = predicted_logits.shape
    prediction = prediction.contiguous()
    if prediction.dtype not in (torch.float16, torch.float32):
        prediction = prediction.float()
    c2, td, th, tw = prediction.shape
    if c2 != cch or (n_predictions is not None and tuple(n_predictions.shape) != (vd, vh, vw)):
        raise RehrError("sw_accumulate: shape mismatch")
    od, oh, ow = (int(v) for v in origin)
    if od < 0 or oh < 0 or ow < 0 or od + td > vd or oh + th > vh or ow + tw > vw:
        raise RehrError("sw_accumulate: tile outside the volume")
    g = None
    if isinstance(gaussian, torch.Tensor):
        g = gaussian.to(torch.float16).contiguous()
        if tuple(g.shape) != (td, th, tw):
            raise RehrError("sw_accumulate: gaussian / tile shape mismatch")
    elif gaussian != 1:
        raise RehrError("gaussian must be a tensor or the literal 1 (utils/seg_utils.py:264-265)")
    check(lib().rehr_sw_accumulate(ptr(predicted_logits), ptr(n_predictions), ptr(prediction), int(prediction.dtype == torch.float32),
                                   ptr(g), cch, vd, vh, vw, td, th, tw, od, oh, ow, stream_ptr()), "sw_accumulate")
    F_._count()


def sw_finalize(predicted_logits: torch.Tensor, n_predictions: torch.Tensor) -> bool:
    """logits /= n (fp16) and the reference's inf check (utils/seg_utils.py:278-283).  Returns True if an inf was produced."""
    flag = torch.zeros((1,), dtype=torch.int32, device=predicted_logits.device)
    cch = predicted_logits.shape[0]
    check(lib().rehr_sw_finalize(ptr(predicted_logits), ptr(n_predictions), cch, n_predictions.numel(), ptr(flag), stream_ptr()),
          "sw_finalize")
    F_._count()
    return bool(flag.item())


INF_MESSAGE = ('Encountered inf in predicted array. Aborting... If this problem persists, '
               'reduce value_scaling_factor in compute_gaussian or increase the dtype of '
               'predicted_logits to fp32')


def _internal_predict_sliding_window_return_logits(data: torch.Tensor, slicers, network, do_on_device=True, out_idx=None,
                                                   slice_seperation=1, patch_size=[14, 320, 384], use_gaussian=False,
                                                   deep_supervision=True, accum_dtype: Optional[torch.dtype] = torch.float16,
                                                   tile_filter: Optional[Callable[[int], bool]] = None, finalize: bool = True,
                                                   cuda_graph: bool = True):
    """utils/seg_utils.py:240-287.  `data` [C_in, X, Y, Z]; returns fp16 logits [2, X*slice_seperation, Y, Z].
    `tile_filter(i)` / `finalize=False` are the hooks the sharded driver uses (returns (logits, n_predictions) then).
    `cuda_graph`: when gradients are off and `network` is an nn.Module, its tile forward is captured once in a CUDA graph
    and replayed for the 8 x n_tiles identically shaped calls (rehrseg_b200/graphs.py)."""
    if not do_on_device:
        raise RehrError("rehrseg_b200 blends on the GPU only (do_on_device=True); there is no CPU path")
    dev = data.device if data.is_cuda else torch.device("cuda", torch.cuda.current_device())
    data = data.to(dev)
    predicted_logits = torch.zeros((2, data.shape[1] * slice_seperation, data.shape[2], data.shape[3]), dtype=torch.half, device=dev)
    n_predictions = torch.zeros((data.shape[1] * slice_seperation, data.shape[2], data.shape[3]), dtype=torch.half, device=dev)
    gaussian = importance_map(patch_size, 1. / 8, 10, device=dev) if use_gaussian else 1
    if out_idx == 0 and isinstance(network, torch.nn.Module) and hasattr(network, "sr_head") and hasattr(network, "upscale"):
        from .seg_model import LRHeadOnly, _EngineForward, SegModel
        if isinstance(network, (SegModel, _EngineForward)):
            # output 1 (the x`upscale` SR head) is never read on this path: do not compute it
            wrapper = network.__dict__.get("_rehr_lr_only")  # kept on the model (not a registered sub-module) so the
            if wrapper is None:                               # captured CUDA graph of the view is reused across calls
                wrapper = LRHeadOnly(network)
                object.__setattr__(network, "_rehr_lr_only", wrapper)
            network = wrapper
    pair_model = None
    if len(slicers) > 0 and isinstance(network, torch.nn.Module) and not torch.is_grad_enabled():
        # mirror variants several per forward (python-launched or graph-replayed alike, so the two give identical bits): the deep
        # (<= 16^3) layers and the tails of the persistent conv kernels fill the machine better with more samples, and the host
        # issues fewer launches
        if SW_PAIR_VARIANTS and SW_VARIANT_GROUP > 1:
            pair_model = network
        if cuda_graph:
            from .graphs import graphed
            if pair_model is not None:
                pair_model = graphed(network, data[slicers[0]][None].repeat(SW_VARIANT_GROUP, 1, 1, 1, 1))
            else:
                network = graphed(network, data[slicers[0]][None])
    for i, sl in enumerate(slicers):
        if tile_filter is not None and not tile_filter(i):
            continue
        workon = data[sl][None]
        prediction = _internal_maybe_mirror_and_predict(network, workon, out_idx, deep_supervision, i == len(slicers) - 1,
                                                        accum_dtype=accum_dtype, pair_model=pair_model, group=SW_VARIANT_GROUP)
        prediction = prediction[0]
        origin = (sl[1].start * slice_seperation, sl[2].start, sl[3].start)
        sw_accumulate(predicted_logits, n_predictions, prediction, gaussian, origin)
    if not finalize:
        return predicted_logits, n_predictions
    if sw_finalize(predicted_logits, n_predictions):
        raise RuntimeError(INF_MESSAGE)
    if use_gaussian:
        compute_gaussian.cache_clear()
    return predicted_logits


def predict_sliding_window_sharded(data: torch.Tensor, slicers, network, group=None, out_idx=None, slice_seperation=1,
                                   patch_size=[14, 320, 384], use_gaussian=True, deep_supervision=True,
                                   accumulate_fn=None, shard_mirrors: bool = True, tiles_per_allreduce: int = 32):
    """N-GPU sliding window (one process per GPU).  The independent units are the (tile, mirror variant) forwards
    (27 x 8 = 216 for config 3): unit u = 8 * tile + variant runs on rank u % world, so the load is balanced even when the
    tile count is not a multiple of the rank count.  Every rank sums the variants it ran per tile (fp32); ONE all-reduce per
    chunk of `tiles_per_allreduce` tiles completes the per-tile sums on every rank; then every rank blends ALL tiles locally in
    the reference's tile order with the reference's fp16 arithmetic (utils/seg_utils.py:256-283) and finalises.  The result is
    the same on every rank and differs from the single-GPU path only by the fp32 summation order of the 8 variants of a tile.
    `shard_mirrors=False` deals whole tiles instead (each rank blends its tiles, fp32 all-reduce of the volume buffers);
    `accumulate_fn` lets the CPU/gloo test substitute the blend of that variant."""
    import torch.distributed as dist
    rank = dist.get_rank(group) if dist.is_initialized() else 0
    world = dist.get_world_size(group) if dist.is_initialized() else 1
    if accumulate_fn is None and (world == 1 or shard_mirrors):
        if world == 1:
            return _internal_predict_sliding_window_return_logits(data, slicers, network, True, out_idx, slice_seperation, patch_size,
                                                                  use_gaussian, deep_supervision)
        return _blend_mirror_units(data, slicers, network, out_idx, slice_seperation, patch_size, use_gaussian, deep_supervision,
                                   rank, world, group, tiles_per_allreduce)
    if accumulate_fn is not None:
        logits, npred = accumulate_fn(data, [(i, sl) for i, sl in enumerate(slicers) if i % world == rank])
    else:
        logits, npred = _internal_predict_sliding_window_return_logits(
            data, slicers, network, True, out_idx, slice_seperation, patch_size, use_gaussian, deep_supervision,
            tile_filter=lambda i: i % world == rank, finalize=False)
    if world > 1:
        l32, n32 = logits.float(), npred.float()
        dist.all_reduce(l32, group=group)
        dist.all_reduce(n32, group=group)
        out = l32 / n32
    else:
        out = logits.float() / npred.float()
    if torch.any(torch.isinf(out)):
        raise RuntimeError(INF_MESSAGE)
    return out.to(torch.float16)


def _blend_mirror_units(data, slicers, network, out_idx, slice_seperation, patch_size, use_gaussian, deep_supervision, rank, world,
                        group, tiles_per_allreduce: int = 32):
    """(tile, mirror variant) units dealt round-robin; per-tile variant sums all-reduced in chunks; local blend of all tiles."""
    import torch.distributed as dist
    group_net = None
    if not torch.is_grad_enabled() and isinstance(network, torch.nn.Module):
        if out_idx == 0 and hasattr(network, "sr_head") and hasattr(network, "upscale"):
            from .seg_model import LRHeadOnly, _EngineForward, SegModel
            if isinstance(network, (SegModel, _EngineForward)):
                wrapper = network.__dict__.get("_rehr_lr_only")
                if wrapper is None:
                    wrapper = LRHeadOnly(network)
                    object.__setattr__(network, "_rehr_lr_only", wrapper)
                network = wrapper
        from .graphs import graphed
        if len(slicers) > 0:
            module = network
            network = graphed(module, data[slicers[0]][None])
            if SW_PAIR_VARIANTS and SW_SHARD_GROUP > 1:
                # this rank's units several per forward (as in the single-GPU path); a remainder runs one by one
                group_net = graphed(module, data[slicers[0]][None].repeat(SW_SHARD_GROUP, 1, 1, 1, 1))
    dev = data.device
    logits = torch.zeros((2, data.shape[1] * slice_seperation, data.shape[2], data.shape[3]), dtype=torch.half, device=dev)
    npred = torch.zeros((data.shape[1] * slice_seperation, data.shape[2], data.shape[3]), dtype=torch.half, device=dev)
    gaussian = importance_map(patch_size, 1. / 8, 10, device=dev) if use_gaussian else 1
    variants = [()] + mirror_axes_combinations()
    nv = len(variants)
    for c0 in range(0, len(slicers), tiles_per_allreduce):
        chunk = list(range(c0, min(len(slicers), c0 + tiles_per_allreduce)))
        sums = None
        mine = [(k, i, v) for k, i in enumerate(chunk) for v in range(nv) if (i * nv + v) % world == rank]
        pos = 0
        while pos < len(mine):
            take = SW_SHARD_GROUP if (group_net is not None and len(mine) - pos >= SW_SHARD_GROUP) else 1
            units = mine[pos:pos + take]
            pos += take
            xs = []
            for _k, i, v in units:
                workon = data[slicers[i]][None]
                xs.append(torch.flip(workon, variants[v]) if variants[v] else workon)
            out = _select((group_net if take > 1 else network)(torch.cat(xs, dim=0) if take > 1 else xs[0]), out_idx, deep_supervision)
            for j, (k, _i, v) in enumerate(units):          # same accumulation order as one unit per forward
                p = out[j:j + 1]
                p = (torch.flip(p, variants[v]) if variants[v] else p).float()
                if sums is None:
                    sums = torch.zeros((len(chunk), *p.shape[1:]), dtype=torch.float32, device=dev)
                sums[k] += p[0]
        if sums is None:   # fewer units than ranks in this chunk: this rank still takes part in the collective
            t = [s.stop - s.start for s in slicers[chunk[0]][1:]]
            sums = torch.zeros((len(chunk), logits.shape[0], t[0] * slice_seperation, t[1], t[2]), dtype=torch.float32, device=dev)
        dist.all_reduce(sums, group=group)
        for k, i in enumerate(chunk):
            sl = slicers[i]
            pred = (sums[k] / nv).to(torch.float16)
            sw_accumulate(logits, npred, pred, gaussian, (sl[1].start * slice_seperation, sl[2].start, sl[3].start))
    if sw_finalize(logits, npred):
        raise RuntimeError(INF_MESSAGE)
    return logits


def sliding_window_segment(model, lr_data, patch_size, slice_separation=1, out_idx=0, use_gaussian=True):
    """The tensor part of `evaluate_case` (utils/seg_utils.py:741-763): zero-pad to at least one tile, tile, blend,
    crop the padding, softmax, argmax.  `lr_data` [1, X, Y, Z] float tensor (already z-scored).  Returns uint8 labels."""
    lr_data, slicer_revert = pad_nd_image(lr_data, patch_size, 'constant', {'value': 0}, True)
    was_ds = getattr(getattr(model, "decoder", None), "deep_supervision", False)
    with torch.no_grad():
        slicers = _internal_get_sliding_window_slicers(lr_data.shape[1:], patch_size=patch_size)
        predicted_logits = _internal_predict_sliding_window_return_logits(
            lr_data, slicers, model, out_idx=out_idx, slice_seperation=1, patch_size=patch_size, use_gaussian=use_gaussian,
            deep_supervision=was_ds)
    prediction = predicted_logits[tuple([slice(None), *slicer_revert[1:]])]
    prediction = torch.softmax(prediction.float(), 0).argmax(0)
    return prediction.to(torch.uint8)


def calculate_dice(prediction, ground_truth, smooth: float = 1e-5) -> float:
    """utils/seg_utils.py:730-734: (2 |P n G| + smooth) / (|P| + |G| + smooth) on flattened label arrays (tensors or numpy)."""
    if isinstance(prediction, torch.Tensor) or isinstance(ground_truth, torch.Tensor):
        p = torch.as_tensor(prediction).flatten().double()
        g = torch.as_tensor(ground_truth).to(p.device).flatten().double()
        return float((2.0 * (p * g).sum() + smooth) / (p.sum() + g.sum() + smooth))
    p, g = np.asarray(prediction).flatten(), np.asarray(ground_truth).flatten()
    return float((2.0 * np.sum(p * g) + smooth) / (np.sum(p) + np.sum(g) + smooth))


def evaluate_case_tensors(model, lr_data: torch.Tensor, lr_label: Optional[torch.Tensor], slice_separation: int, patch_size,
                          get_HR_results: bool = False):
    """The tensor part of `evaluate_case` (utils/seg_utils.py:736-784) -- everything between `preprocess_image` (file IO, out of
    scope) and the return: pad to at least one tile, sliding window over output 0 (Gaussian blend, mirror TTA), crop the padding,
    softmax / argmax -> uint8 LR labels, Dice against `lr_label`; with `get_HR_results` a second sliding window over output 1
    (the x`slice_separation` SR head; no Gaussian, as the reference's call leaves `use_gaussian` at its default).
    `lr_data` [1, X, Y, Z] (already z-scored), `lr_label` [1, X, Y, Z] or None.  Returns (prediction_lr, prediction_hr, dice_lr)."""
    model.eval()
    patch_size = list(patch_size)
    dev = lr_data.device if lr_data.is_cuda else torch.device("cuda", torch.cuda.current_device())
    lr_data = lr_data.to(dev)
    lr_data, slicer_revert = pad_nd_image(lr_data, patch_size, 'constant', {'value': 0}, True)
    with torch.no_grad():
        slicers = _internal_get_sliding_window_slicers(lr_data.shape[1:], patch_size=patch_size)
        logits = _internal_predict_sliding_window_return_logits(lr_data, slicers, model, True, 0, 1, patch_size, use_gaussian=True,
                                                                deep_supervision=False)
        prediction = logits[tuple([slice(None), *slicer_revert[1:]])]
        prediction_lr = torch.softmax(prediction.float(), dim=0).argmax(0).to(torch.uint8)
        dice_lr = calculate_dice(prediction_lr, lr_label.squeeze(0).to(torch.uint8)) if lr_label is not None else None
        if get_HR_results:
            hr_patch = [patch_size[0] * slice_separation, patch_size[1], patch_size[2]]
            hr_logits = _internal_predict_sliding_window_return_logits(lr_data, slicers, model, True, 1, slice_separation, hr_patch)
            prediction_hr = torch.argmax(hr_logits, dim=0).to(torch.uint8)
        else:
            prediction_hr = prediction_lr
    return prediction_lr, prediction_hr, dice_lr
