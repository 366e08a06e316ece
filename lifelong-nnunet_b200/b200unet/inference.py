"""Sliding-window inference on the CUDA path: what the reference's evaluation / inference sweep
(inference/predict.py:117-401, evaluation/evaluator.py:70-330, `_perform_validation` MultiHead:678-901) obtains from nnunet's
``SegmentationNetwork.predict_3D`` -> ``_internal_predict_3D_3Dconv_tiled`` (un-vendored nnunet@77bc485):

  pad the case to at least the patch size (constant zeros, centred) -> patch origins from ``_compute_steps_for_sliding_window``
  (step_size 0.5) -> per patch: network forward (optionally averaged over mirrored copies), softmax, multiplication with the
  Gaussian importance map (sigma = patch / 8), accumulation -> division by the accumulated weights -> argmax -> crop.

The network forward is the plan's forward (`Generic_UNet.__call__`, no_grad); softmax / un-mirroring / Gaussian weighting /
accumulation are ONE fused launch per patch (b2_sliding_accumulate), normalisation + argmax one launch (b2_sliding_finalize).
"""
import ctypes as C

import numpy as np
import torch

from . import _lib


def compute_steps_for_sliding_window(patch_size, image_size, step_size):
    """nnunet SegmentationNetwork._compute_steps_for_sliding_window"""
    assert all(i >= j for i, j in zip(image_size, patch_size)), "image size must be as large or larger than patch_size"
    assert 0 < step_size <= 1, 'step_size must be larger than 0 and smaller or equal to 1'
    target = [i * step_size for i in patch_size]
    num_steps = [int(np.ceil((i - k) / j)) + 1 for i, j, k in zip(image_size, target, patch_size)]
    steps = []
    for dim in range(len(patch_size)):
        max_step = image_size[dim] - patch_size[dim]
        actual = max_step / (num_steps[dim] - 1) if num_steps[dim] > 1 else 99999999999
        steps.append([int(np.round(actual * i)) for i in range(num_steps[dim])])
    return steps


def get_gaussian(patch_size, sigma_scale=1. / 8):
    """nnunet SegmentationNetwork._get_gaussian: a unit impulse at the patch centre blurred with sigma = patch * sigma_scale,
    scaled to max 1, zeros replaced by the smallest positive value"""
    from scipy.ndimage import gaussian_filter
    tmp = np.zeros(patch_size)
    tmp[tuple(i // 2 for i in patch_size)] = 1
    g = gaussian_filter(tmp, [i * sigma_scale for i in patch_size], 0, mode='constant', cval=0)
    g = g / np.max(g) * 1
    g = g.astype(np.float32)
    g[g == 0] = np.min(g[g != 0])
    return g


def _pad_to_patch(x, patch_size):
    """batchgenerators pad_nd_image(mode='constant', kwargs {'constant_values': 0}, return_slicer=True): centred padding to at
    least the patch size; returns the padded array and the slicer that crops the result back"""
    shape = x.shape[1:]
    new = [max(s, p) for s, p in zip(shape, patch_size)]
    diff = [n - s for n, s in zip(new, shape)]
    below = [d // 2 for d in diff]
    above = [d - b for d, b in zip(diff, below)]
    if any(diff):
        x = torch.nn.functional.pad(x, (below[2], above[2], below[1], above[1], below[0], above[0]))
    slicer = tuple(slice(b, b + s) for b, s in zip(below, shape))
    return x, slicer


@torch.no_grad()
def predict_3D(network, x, patch_size, do_mirroring=True, mirror_axes=(0, 1, 2), step_size=0.5, use_gaussian=True):
    """x: (C, D, H, W) tensor or ndarray.  Returns (segmentation (D, H, W) int64, class probabilities (ncls, D, H, W) fp32)
    like nnunet's predict_3D(..., use_sliding_window=True, all_in_gpu=True)."""
    lib = _lib.load()
    dev = next(network.parameters()).device
    x = torch.as_tensor(x, dtype=torch.float32).to(dev)
    assert x.dim() == 4, "data must have shape (c, x, y, z)"
    patch_size = tuple(int(p) for p in patch_size)
    data, slicer = _pad_to_patch(x, patch_size)
    Dd, Hh, Ww = (int(s) for s in data.shape[1:])
    steps = compute_steps_for_sliding_window(patch_size, (Dd, Hh, Ww), step_size)
    ncls = network.num_classes
    if ncls > 8:
        raise NotImplementedError("b2_sliding_accumulate handles up to 8 classes")
    gauss = torch.from_numpy(get_gaussian(patch_size)).to(dev) if use_gaussian and sum(len(s) for s in steps) > 3 else None
    agg = torch.zeros((ncls, Dd, Hh, Ww), dtype=torch.float32, device=dev)
    wsum = torch.zeros((Dd, Hh, Ww), dtype=torch.float32, device=dev)
    flips = [()]
    if do_mirroring:
        flips = [tuple(a for a in mirror_axes if (m >> mirror_axes.index(a)) & 1) for m in range(2 ** len(mirror_axes))]
    scale = 1.0 / len(flips)
    st = C.c_void_p(torch.cuda.current_stream(dev).cuda_stream)
    was_ds, network.do_ds = network.do_ds, False
    try:
        for z0 in steps[0]:
            for y0 in steps[1]:
                for x0 in steps[2]:
                    patch = data[None, :, z0:z0 + patch_size[0], y0:y0 + patch_size[1], x0:x0 + patch_size[2]]
                    for k, axes in enumerate(flips):
                        inp = torch.flip(patch, [a + 2 for a in axes]) if axes else patch
                        logits = network(inp.contiguous())
                        mask = sum(1 << a for a in axes)
                        _lib.check(lib.b2_sliding_accumulate(C.c_void_p(logits.data_ptr()), ncls, *patch_size,
                                                             None if gauss is None else C.c_void_p(gauss.data_ptr()),
                                                             C.c_void_p(agg.data_ptr()), C.c_void_p(wsum.data_ptr()), Dd, Hh, Ww,
                                                             int(z0), int(y0), int(x0), mask, scale, 1 if k == 0 else 0, st))
    finally:
        network.do_ds = was_ds
    seg = torch.empty((Dd, Hh, Ww), dtype=torch.int32, device=dev)
    _lib.check(lib.b2_sliding_finalize(C.c_void_p(agg.data_ptr()), C.c_void_p(wsum.data_ptr()), ncls, Dd * Hh * Ww,
                                       C.c_void_p(seg.data_ptr()), st))
    return seg[slicer].long(), agg[(slice(None),) + slicer]
