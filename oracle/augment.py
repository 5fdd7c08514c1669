"""CPU oracle of the stage-2 spatial augmentation: `augment_spatial` / `MySpatialTransform`
(utils/seg_utils.py:378-631) in the configuration `get_training_transforms` builds for the stage-2 dataset
(utils/seg_utils.py:652-676, utils/train_set.py:64-85): dummy-2D (the volume's slices become channels), no elastic deformation,
rotation about x with p = 0.2, isotropic scaling in (0.7, 1.4) with p = 0.2, no random crop, order-3 spline interpolation of the
image (and of the uncertainty map), order-1 per-label interpolation of the segmentations, constant borders.
TEST INFRASTRUCTURE: imported only by tests/ and oracle/make_golden.py.

Third-party pieces (absent from /root/reference and from the image, restated from their published source -- "parity unpinned" for
these definitions, anchored on the reference's call sites):
  batchgenerators 0.25 (requirements.txt:5) `augmentations.utils`: create_zero_centered_coordinate_mesh, rotate_coords_2d,
    scale_coords, interpolate_img;  nnunetv2 2.3.1 `transforms_for_dummy_2d`: Convert3DTo2DTransform / Convert2DTo3DTransform.
scipy.ndimage.map_coordinates IS installed (here and on the GPU box) and is called directly.
The fixture tests/golden/spatial_aug.npz is produced by the reference's OWN `augment_spatial` with these helpers injected
(oracle/make_golden.py); `augment_spatial_2d` below must reproduce it bit for bit under the same `np.random` seed."""
from __future__ import annotations

import numpy as np
from scipy.ndimage import map_coordinates


# ---- batchgenerators.augmentations.utils (restated) -----------------------------------------------------------------------
def create_zero_centered_coordinate_mesh(shape):
    tmp = tuple([np.arange(i) for i in shape])
    coords = np.array(np.meshgrid(*tmp, indexing="ij")).astype(float)
    for d in range(len(shape)):
        coords[d] -= ((np.array(shape).astype(float) - 1) / 2.)[d]
    return coords


def create_matrix_rotation_2d(angle, matrix=None):
    rotation = np.array([[np.cos(angle), -np.sin(angle)], [np.sin(angle), np.cos(angle)]])
    if matrix is None:
        return rotation
    return np.dot(matrix, rotation)


def rotate_coords_2d(coords, angle):
    rot_matrix = create_matrix_rotation_2d(angle)
    return np.dot(coords.reshape(len(coords), -1).transpose(), rot_matrix).transpose().reshape(coords.shape)


def scale_coords(coords, scale):
    if isinstance(scale, (tuple, list, np.ndarray)):
        assert len(scale) == len(coords)
        for i in range(len(scale)):
            coords[i] *= scale[i]
    else:
        coords *= scale
    return coords


def interpolate_img(img, coords, order=3, mode="nearest", cval=0.0, is_seg=False):
    if is_seg and order != 0:
        unique_labels = np.unique(img)
        result = np.zeros(coords.shape[1:], img.dtype)
        for c in unique_labels:
            res_new = map_coordinates((img == c).astype(float), coords, order=order, mode=mode, cval=cval)
            result[res_new >= 0.5] = c
        return result
    return map_coordinates(img.astype(float), coords, order=order, mode=mode, cval=cval).astype(img.dtype)


def rotate_coords_3d(*_a, **_k):          # (imported by the reference module; the dummy-2D configuration never calls them)
    raise NotImplementedError


def elastic_deform_coordinates(*_a, **_k):
    raise NotImplementedError


# ---- nnunetv2 transforms_for_dummy_2d (restated) -----------------------------------------------------------------------------
def convert_3d_to_2d(arr: np.ndarray):
    """[b, c, z, x, y] -> ([b, c*z, x, y], original shape)"""
    shp = arr.shape
    return arr.reshape((shp[0], shp[1] * shp[2], shp[3], shp[4])), shp


def convert_2d_to_3d(arr: np.ndarray, shp):
    return arr.reshape((shp[0], shp[1], shp[2], arr.shape[-2], arr.shape[-1]))


# ---- utils/seg_utils.py:378-458, dim == 2, do_elastic_deform = False -------------------------------------------------------------
def draw_affine_2d(rng, do_rotation=True, angle_x=(-np.pi, np.pi), do_scale=True, scale=(0.7, 1.4), p_scale_per_sample=0.2,
                   p_rot_per_sample=0.2, p_rot_per_axis=1.0):
    """The random decisions of ONE sample in the reference's order (utils/seg_utils.py:409-446): (angle or None, scale or None)."""
    a_x = None
    if do_rotation and rng.uniform() < p_rot_per_sample:
        a_x = rng.uniform(angle_x[0], angle_x[1]) if rng.uniform() <= p_rot_per_axis else 0
    sc = None
    if do_scale and rng.uniform() < p_scale_per_sample:
        if rng.random() < 0.5 and scale[0] < 1:
            sc = rng.uniform(scale[0], 1)
        else:
            sc = rng.uniform(max(scale[0], 1), scale[1])
    return a_x, sc


def augment_spatial_2d(data, seg_list, patch_size, do_rotation=True, angle_x=(-np.pi, np.pi), do_scale=True, scale=(0.7, 1.4),
                       border_mode_data="constant", border_cval_data=0, order_data=3, border_mode_seg="constant", border_cval_seg=-1,
                       order_seg=1, p_scale_per_sample=0.2, p_rot_per_sample=0.2, p_rot_per_axis=1.0, enable_uncertainty=False,
                       rng=np.random):
    """`augment_spatial` for 2-D batches [b, c, x, y] with the stage-2 settings (no elastic deformation, no random crop)."""
    seg_result = None
    if seg_list is not None:
        seg_result = [np.zeros((x.shape[0], x.shape[1], patch_size[0], patch_size[1]), dtype=np.float32) for x in seg_list]
    data_result = np.zeros((data.shape[0], data.shape[1], patch_size[0], patch_size[1]), dtype=np.float32)
    for sample_id in range(data.shape[0]):
        coords = create_zero_centered_coordinate_mesh(patch_size)
        a_x, sc = draw_affine_2d(rng, do_rotation, angle_x, do_scale, scale, p_scale_per_sample, p_rot_per_sample, p_rot_per_axis)
        if a_x is not None:
            coords = rotate_coords_2d(coords, a_x)
        if sc is not None:
            coords = scale_coords(coords, sc)
        for d in range(2):
            coords[d] += data.shape[d + 2] / 2. - 0.5
        for channel_id in range(data.shape[1]):
            data_result[sample_id, channel_id] = interpolate_img(data[sample_id, channel_id], coords, order_data, border_mode_data,
                                                                 cval=border_cval_data)
        if seg_list is not None:
            for i, seg in enumerate(seg_list):
                for channel_id in range(seg.shape[1]):
                    if i == len(seg_list) - 1 and enable_uncertainty:
                        seg_result[i][sample_id, channel_id] = interpolate_img(seg[sample_id, channel_id], coords, order_data,
                                                                               border_mode_data, cval=border_cval_data, is_seg=False)
                    else:
                        seg_result[i][sample_id, channel_id] = interpolate_img(seg[sample_id, channel_id], coords, order_seg,
                                                                               border_mode_seg, cval=border_cval_seg, is_seg=True)
    return data_result, seg_result


def spatial_transform_dummy_2d(data_dict: dict, patch_size_zxy, keys=("seg", "seg_sr", "uncertainty"), enable_uncertainty=True,
                               rng=np.random, **aug_kw) -> dict:
    """Convert3DTo2DTransform -> MySpatialTransform(patch_size[1:]) -> Convert2DTo3DTransform on {'data', *keys} (numpy
    [b, c, z, x, y]), utils/seg_utils.py:652-676."""
    shapes = {}
    flat = {}
    for k in ("data", *keys):
        flat[k], shapes[k] = convert_3d_to_2d(data_dict[k])
    d, segs = augment_spatial_2d(flat["data"], [flat[k] for k in keys], tuple(patch_size_zxy[1:]), enable_uncertainty=enable_uncertainty,
                                 rng=rng, **aug_kw)
    out = {"data": convert_2d_to_3d(d, shapes["data"])}
    for k, s in zip(keys, segs):
        out[k] = convert_2d_to_3d(s, shapes[k])
    return out


# ---- TrainSetMultipleSegSREfficient.__getitem__ (utils/train_set.py:102-159) -----------------------------------------------------
def zscore_normalization_np(image: np.ndarray) -> np.ndarray:
    """utils/seg_utils.py:149-155 (numpy branch)."""
    image = image.astype(np.float32, copy=False)
    mean = image.mean()
    std = image.std()
    image -= mean
    image /= (max(std, 1e-8))
    return image


def _target_pad_const(img, target_dims):
    from .volume import get_pads
    pads = tuple(get_pads(t, d) for t, d in zip(target_dims, img.shape))
    return np.pad(img, pads, mode="constant")


def stage2_sample(img, label, uncertainty, patch_size, separation: int, norm=True, random_flip=True, use_uncertainty=True,
                  transform=None, rng_py=None):
    """The reference's `__getitem__` statement by statement; `transform(**dict)` stands for `self.train_transform` (the spatial part
    of it is `spatial_transform_dummy_2d` + a tensor conversion; the intensity transforms are third-party batchgenerators classes).
    img / label / uncertainty: [X, Y, Z] arrays; patch_size = (x, y, z_lr)."""
    import random as _random
    import torch
    rng = rng_py or _random
    img = img.copy()
    if norm:
        img = zscore_normalization_np(img)
    ps = patch_size
    x_0 = rng.randint(0, max(img.shape[0] - ps[0], 0))
    y_0 = rng.randint(0, max(img.shape[1] - ps[1], 0))
    z_0 = rng.randint(0, max(img.shape[2] - ps[2] * separation, 0))
    sl = (slice(x_0, x_0 + ps[0]), slice(y_0, y_0 + ps[1]), slice(z_0, z_0 + ps[2] * separation))
    img = img[sl]
    target_shape = [max(s, p) for s, p in zip(img.shape, (ps[0], ps[1], ps[2] * separation))]
    img = _target_pad_const(img, target_shape)
    label = _target_pad_const(label[sl], target_shape)
    if use_uncertainty:
        uncertainty = _target_pad_const(uncertainty[sl], target_shape)
    if random_flip:
        for axis in (0, 1, 2):
            if rng.random() < 0.5:
                img = np.flip(img, axis=axis)
                label = np.flip(label, axis=axis)
                uncertainty = np.flip(uncertainty, axis=axis) if use_uncertainty else None
    img = img[:, :, ::separation]
    label_lr = label[:, :, ::separation]
    img = img.copy().transpose(2, 1, 0)[None, None, ...]
    label = label.copy().transpose(2, 1, 0)[None, None, ...]
    label_lr = label_lr.copy().transpose(2, 1, 0)[None, None, ...]
    if use_uncertainty:
        uncertainty_lr = uncertainty[:, :, ::separation]
        uncertainty_lr = uncertainty_lr.copy().transpose(2, 1, 0)[None, None, ...]
        uncertainty_lr = 1 - uncertainty_lr / 255. * 0.99
        out_data = transform(**{"data": img.astype("float32"), "seg": label_lr, "seg_sr": label, "uncertainty": uncertainty_lr})
        uncertainty_lr = out_data["uncertainty"].squeeze(0)
    else:
        out_data = transform(**{"data": img.astype("float32"), "seg": label_lr, "seg_sr": label})
        uncertainty_lr = 0
    return out_data["data"].squeeze(0), out_data["seg"].squeeze(0), out_data["seg_sr"].squeeze(0), uncertainty_lr


def spatial_only_transform(patch_size_zyx, enable_uncertainty=True, rng=np.random, **aug_kw):
    """`get_training_transforms` (utils/seg_utils.py:632-727) reduced to its in-tree spatial part + NumpyToTensor('float')."""
    import torch
    keys = ("seg", "seg_sr", "uncertainty") if enable_uncertainty else ("seg", "seg_sr")

    def run(**dd):
        out = spatial_transform_dummy_2d(dd, patch_size_zyx, keys=keys, enable_uncertainty=enable_uncertainty, rng=rng, **aug_kw)
        return {k: torch.from_numpy(np.ascontiguousarray(v)).float() for k, v in out.items()}

    return run
