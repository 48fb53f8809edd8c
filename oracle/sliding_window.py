"""Restatement of nnU-Net v1 ``SegmentationNetwork.predict_3D`` (tiled, Conv3d path) — SURVEY.md Appendix A.

UPSTREAM: ``nnunet/network_architecture/neural_network.py`` of the KAIST BraTS-2021 fork
(``Brats21_KAIST_MRI_Lab/nnunet``; git-ignored by the reference, no pinned version, absent from /root/reference).
Anchored on the reference call site run_brats2021_inference_singlethread.py:97-106 (do_mirroring=True,
mirror_axes from the trainer, use_sliding_window=True, step_size=0.5, use_gaussian=True, all_in_gpu=False,
mixed_precision=True -> fp32 on CPU because autocast is CUDA-only).  Parity for this sub-part is unpinned by
reference tests; known-answer facts live in tests/golden/sliding_window.json.
"""
import numpy as np
import torch
from scipy.ndimage import gaussian_filter


def compute_steps_for_sliding_window(patch_size, image_size, step_size):
    """App. A.3 `_compute_steps_for_sliding_window`."""
    assert all(i >= j for i, j in zip(image_size, patch_size)), "image size must be >= patch size"
    assert 0 < step_size <= 1
    target = [i * step_size for i in patch_size]
    num_steps = [int(np.ceil((i - k) / j)) + 1 for i, j, k in zip(image_size, target, patch_size)]
    steps = []
    for dim in range(len(patch_size)):
        max_step = image_size[dim] - patch_size[dim]
        actual = max_step / (num_steps[dim] - 1) if num_steps[dim] > 1 else 99999999999
        steps.append([int(np.round(actual * i)) for i in range(num_steps[dim])])
    return steps


def get_gaussian(patch_size, sigma_scale=1.0 / 8):
    """App. A.4 `_get_gaussian`: centred delta -> scipy gaussian_filter -> /max -> fp32 -> zeros := min non-zero."""
    tmp = np.zeros(patch_size)
    center = [i // 2 for i in patch_size]
    sigmas = [i * sigma_scale for i in patch_size]
    tmp[tuple(center)] = 1
    g = gaussian_filter(tmp, sigmas, 0, mode="constant", cval=0)
    g = g / np.max(g) * 1
    g = g.astype(np.float32)
    g[g == 0] = np.min(g[g != 0])
    return g


def pad_nd_image(image, new_shape, mode="constant", kwargs=None, shape_must_be_divisible_by=None):
    """batchgenerators `pad_nd_image` restricted to what predict_3D uses: symmetric pad of the trailing dims up to
    new_shape (floor on the low side) and, when `shape_must_be_divisible_by` is given (the non-tiled path), further up
    to the next multiple per axis; returns (padded, slicer)."""
    kwargs = kwargs or {"constant_values": 0}
    old = np.array(image.shape[-len(new_shape):])
    new = np.array([max(a, b) for a, b in zip(new_shape, old)])
    if shape_must_be_divisible_by is not None:
        div = np.array(shape_must_be_divisible_by)
        new = np.array([n if n % d == 0 else n + d - n % d for n, d in zip(new, div)])
    diff = new - old
    below = diff // 2
    above = diff // 2 + diff % 2
    pad = [[0, 0]] * (image.ndim - len(new_shape)) + [list(p) for p in zip(below, above)]
    if any(p[0] or p[1] for p in pad):
        res = np.pad(image, pad, mode, **kwargs)
    else:
        res = image
    pad = np.array(pad)
    slicer = tuple(slice(int(pad[i][0]), int(res.shape[i] - pad[i][1])) for i in range(image.ndim))
    return res, slicer


MIRROR_FLIPS = [(), (4,), (3,), (4, 3), (2,), (4, 2), (3, 2), (4, 3, 2)]  # App. A.6, tensor dims of (1,C,z,y,x)


def mirror_and_predict(forward_fn, nonlin, x, mirror_axes, do_mirroring, mult, num_classes):
    """App. A.6 `_internal_maybe_mirror_and_pred_3D` on a (1,C,z,y,x) fp32 tensor."""
    result = torch.zeros([1, num_classes] + list(x.shape[2:]), dtype=torch.float32)
    if do_mirroring:
        mirror_idx, num_results = 8, 2 ** len(mirror_axes)
    else:
        mirror_idx, num_results = 1, 1
    for m in range(mirror_idx):
        flips = MIRROR_FLIPS[m]
        if any((f - 2) not in mirror_axes for f in flips):
            continue
        xin = torch.flip(x, flips) if flips else x
        pred = nonlin(forward_fn(xin))
        if flips:
            pred = torch.flip(pred, flips)
        result += 1.0 / num_results * pred
    if mult is not None:
        result[:, :] *= mult
    return result


def predict_3d_tiled(forward_fn, nonlin, x, num_classes, patch_size, do_mirroring=True, mirror_axes=(0, 1, 2),
                     step_size=0.5, use_gaussian=True, regions_class_order=None, tile_hook=None):
    """App. A.5 `_internal_predict_3D_3Dconv_tiled`, all_in_gpu=False branch.

    x: float32 numpy (C, z, y, x).  Returns (seg, class_probabilities) exactly as predict_3D does:
    seg int64 argmax or float32 ordered-threshold map; probabilities float32 (num_classes, z, y, x).
    """
    assert x.ndim == 4
    data, slicer = pad_nd_image(x, patch_size, "constant", {"constant_values": 0})
    data_shape = data.shape
    steps = compute_steps_for_sliding_window(patch_size, data_shape[1:], step_size)
    num_tiles = len(steps[0]) * len(steps[1]) * len(steps[2])
    if use_gaussian and num_tiles > 1:
        gaussian = get_gaussian(patch_size, 1.0 / 8)
        add = gaussian
        gmult = torch.from_numpy(gaussian)
    else:
        gaussian = None
        add = np.ones(patch_size, dtype=np.float32)
        gmult = None
    agg = np.zeros([num_classes] + list(data_shape[1:]), dtype=np.float32)
    nb = np.zeros([num_classes] + list(data_shape[1:]), dtype=np.float32)
    for xs in steps[0]:
        for ys in steps[1]:
            for zs in steps[2]:
                sl = (slice(xs, xs + patch_size[0]), slice(ys, ys + patch_size[1]), slice(zs, zs + patch_size[2]))
                tile = torch.from_numpy(np.ascontiguousarray(data[(None, slice(None)) + sl]))
                pred = mirror_and_predict(forward_fn, nonlin, tile, mirror_axes, do_mirroring, gmult, num_classes)[0]
                pred = pred.numpy()
                agg[(slice(None),) + sl] += pred
                nb[(slice(None),) + sl] += add
                if tile_hook is not None:
                    tile_hook(xs, ys, zs)
    sl = (slice(None),) + slicer[1:]
    agg = agg[sl]
    nb = nb[sl]
    probs = agg / nb
    seg = decide(probs, regions_class_order)
    return seg, probs


def predict_3d_full(forward_fn, nonlin, x, num_classes, min_size, divisible_by, do_mirroring=True,
                    mirror_axes=(0, 1, 2), regions_class_order=None):
    """nnU-Net v1 `_internal_predict_3D_3Dconv` (predict_3D with use_sliding_window=False; UPSTREAM, restated —
    parity unpinned): pad to >= min_size and to a multiple of `input_shape_must_be_divisible_by`, ONE mirrored forward
    over the whole padded volume, crop back, decide."""
    assert x.ndim == 4
    data, slicer = pad_nd_image(x, min_size, "constant", {"constant_values": 0}, divisible_by)
    pred = mirror_and_predict(forward_fn, nonlin, torch.from_numpy(np.ascontiguousarray(data[None])), mirror_axes,
                              do_mirroring, None, num_classes)[0].numpy()
    probs = pred[(slice(None),) + slicer[1:]]
    return decide(probs, regions_class_order), probs


def decide(probs, regions_class_order):
    """argmax, or the BraTS-regions ordered threshold (App. A.5 / A.7; later regions overwrite earlier ones)."""
    if regions_class_order is None:
        return probs.argmax(0)
    seg = np.zeros(probs.shape[1:], dtype=np.float32)
    for i, c in enumerate(regions_class_order):
        seg[probs[i] > 0.5] = c
    return seg
