"""ORACLE (test infrastructure, CPU / numpy + scipy.ndimage) -- restatement of the patch pipeline the reference obtains from
nnunet@77bc485's DataLoader3D + get_moreDA_augmentation (batchgenerators 0.21), as called at
nnunet_ext/training/network_training/multihead/nnUNetTrainerMultiHead.py:505-511, 904-922.  Both packages are un-vendored and
absent here, so this is written from upstream knowledge ("parity unpinned" at that boundary); what IS pinned is the numerical
core: every interpolation / filter below is done by scipy.ndimage itself (map_coordinates, gaussian_filter, zoom), the library
batchgenerators calls.  Only tests/ may import this module.

`apply_plan` consumes the plain-dict plan of random parameters that b200unet.augment.GPUPatchPipeline.draw_plan produces, so
both sides transform the same crops with the same parameters."""
import numpy as np
from scipy import ndimage as ndi


def rotation_matrix(ax, ay, az):
    """batchgenerators create_matrix_rotation_{x,y,z}_3d chained as in rotate_coords_3d: Rx . Ry . Rz"""
    rx = np.array([[1, 0, 0], [0, np.cos(ax), -np.sin(ax)], [0, np.sin(ax), np.cos(ax)]])
    ry = np.array([[np.cos(ay), 0, np.sin(ay)], [0, 1, 0], [-np.sin(ay), 0, np.cos(ay)]])
    rz = np.array([[np.cos(az), -np.sin(az), 0], [np.sin(az), np.cos(az), 0], [0, 0, 1]])
    return np.dot(np.dot(rx, ry), rz)


def crop_case(case, lb, gen_patch):
    """DataLoader3D.generate_train_batch: valid part of the box, np.pad constant (data 0, segmentation -1)"""
    shape = case.shape[1:]
    ub = [l + g for l, g in zip(lb, gen_patch)]
    vl = [max(0, l) for l in lb]
    vu = [min(s, u) for s, u in zip(shape, ub)]
    part = case[:, vl[0]:vu[0], vl[1]:vu[1], vl[2]:vu[2]]
    pad = [(0, 0)] + [(-min(0, l), max(u - s, 0)) for l, u, s in zip(lb, ub, shape)]
    data = np.pad(part[:-1], pad, "constant", constant_values=0)
    seg = np.pad(part[-1:], pad, "constant", constant_values=-1)
    return data, seg


def spatial(data, seg, sp, patch):
    """batchgenerators augment_spatial (do_elastic_deform False, random_crop False, order_data 3, order_seg 1, border constant
    0 / -1): zero-centred mesh -> rotate -> scale -> + crop centre -> interpolate_img; without a transform: centre crop.
    Returns (data, seg, margin) where margin[z, y, x] = distance (voxels) of the sampling point to the crop border (negative
    outside): the tests skip points within rounding of the border, where fp32 coordinates may fall on the other side."""
    gen = data.shape[1:]
    if sp["angles"] is None and sp["scale"] is None:
        lb = [(g - p) // 2 for g, p in zip(gen, patch)]
        sl = tuple(slice(l, l + p) for l, p in zip(lb, patch))
        return data[(slice(None),) + sl].copy(), seg[(slice(None),) + sl].copy(), np.full(patch, 1e9)
    coords = np.array(np.meshgrid(*[np.arange(p) for p in patch], indexing="ij")).astype(float)
    for d in range(3):
        coords[d] -= (patch[d] - 1) / 2.
    if sp["angles"] is not None:
        coords = np.dot(coords.reshape(3, -1).transpose(), rotation_matrix(*sp["angles"])).transpose().reshape(coords.shape)
    if sp["scale"] is not None:
        coords = coords * sp["scale"]
    for d in range(3):
        coords[d] += gen[d] / 2. - 0.5
    out = np.stack([ndi.map_coordinates(data[c].astype(float), coords, order=3, mode="constant", cval=0.0).astype(np.float32)
                    for c in range(data.shape[0])])
    s = seg[0]
    res = np.zeros(coords.shape[1:], s.dtype)
    for c in np.unique(s):
        m = ndi.map_coordinates((s == c).astype(float), coords, order=1, mode="constant", cval=-1)
        res[m >= 0.5] = c
    margin = np.min([np.minimum(coords[d], gen[d] - 1 - coords[d]) for d in range(3)], 0)
    return out, res[None], margin


def normal_field(seed, start, count):
    """the counter-based generator of aug_pointwise_kernel: splitmix64(seed + (index + 1) * golden) -> Box-Muller"""
    with np.errstate(over="ignore"):
        idx = np.arange(start, start + count, dtype=np.uint64)
        h = np.uint64(seed) + (idx + np.uint64(1)) * np.uint64(0x9E3779B97F4A7C15)
        h = (h ^ (h >> np.uint64(30))) * np.uint64(0xBF58476D1CE4E5B9)
        h = (h ^ (h >> np.uint64(27))) * np.uint64(0x94D049BB133111EB)
        h = h ^ (h >> np.uint64(31))
    u1 = ((h >> np.uint64(40)).astype(np.float32) + np.float32(0.5)) * np.float32(1.0 / 16777216.0)
    u2 = (((h >> np.uint64(16)) & np.uint64(0xFFFFFF)).astype(np.float32) + np.float32(0.5)) * np.float32(1.0 / 16777216.0)
    return (np.sqrt(-2.0 * np.log(u1.astype(np.float64))) * np.cos(2.0 * np.pi * u2.astype(np.float64))).astype(np.float32)


def simulate_low_resolution(x, zoom):
    """batchgenerators augment_linear_downsampling_scipy (per channel): skimage resize(order 0) down to round(shape * zoom) and
    resize(order 3) back, mode 'edge', anti_aliasing off, clip to the input range -- skimage's n-d resize IS
    scipy.ndimage.zoom(grid_mode=True, mode='nearest')"""
    shp = np.array(x.shape)
    target = np.maximum(np.round(shp * zoom).astype(int), 2)
    down = ndi.zoom(x.astype(float), target / shp, order=0, mode="nearest", grid_mode=True)
    up = ndi.zoom(down, shp / target, order=3, mode="nearest", grid_mode=True)
    return np.clip(up, down.min(), down.max()).astype(np.float32)


def gamma(x, g, invert):
    """batchgenerators augment_gamma, per channel, retain_stats=True"""
    if invert:
        x = -x
    mn, sd = x.mean(), x.std()
    minm = x.min()
    rnge = x.max() - minm
    x = np.power((x - minm) / float(rnge + 1e-7), g) * float(rnge + 1e-7) + minm
    x = x - x.mean()
    x = x / (x.std() + 1e-8) * sd
    x = x + mn
    return -x if invert else x


def apply_plan(cases, plan, patch, gen_patch, ds_strides):
    """cases: list of ndarray [C + 1, D, H, W].  Returns (data [B, C, *patch] fp32, [target per scale [B, 1, ...]], margins)"""
    B = len(plan["cases"])
    datas, segs, margins = [], [], []
    for j in range(B):
        d, s = crop_case(cases[plan["cases"][j]], plan["lb"][j], gen_patch)
        d, s, m = spatial(d, s, plan["spatial"][j], patch)
        datas.append(d.astype(np.float32))
        segs.append(s.astype(np.float32))
        margins.append(m)
    Cc = datas[0].shape[0]
    V = int(np.prod(patch))
    for j in range(B):
        d = datas[j]
        if plan["noise"][j] is not None:       # augment_gaussian_noise: the drawn "variance" is used as the standard deviation
            d = d + plan["noise"][j] * normal_field(plan["seed"], j * Cc * V, Cc * V).reshape(d.shape)
        for c in range(Cc):
            x = d[c]
            if plan["blur"][j][c] is not None:
                x = ndi.gaussian_filter(x, plan["blur"][j][c], order=0)
            if plan["brightness"][j][c] is not None:
                x = x * np.float32(plan["brightness"][j][c])
            if plan["contrast"][j][c] is not None:      # augment_contrast, preserve_range=True
                mn, lo, hi = x.mean(), x.min(), x.max()
                x = np.clip((x - mn) * plan["contrast"][j][c] + mn, lo, hi)
            if plan.get("lowres") and plan["lowres"][j][c] is not None:
                x = simulate_low_resolution(x, plan["lowres"][j][c])
            if plan["gamma_inv"][j][c] is not None:
                x = gamma(x, plan["gamma_inv"][j][c], True)
            if plan["gamma"][j][c] is not None:
                x = gamma(x, plan["gamma"][j][c], False)
            d[c] = x
        datas[j] = d
    out_d, out_s = np.stack(datas).astype(np.float32), np.stack(segs)
    margins = np.stack(margins)
    for j in range(B):                                   # MirrorTransform
        for a in range(3):
            if plan["flips"][j] >> a & 1:
                out_d[j] = np.flip(out_d[j], a + 1)
                out_s[j] = np.flip(out_s[j], a + 1)
                margins[j] = np.flip(margins[j], a)
    out_s[out_s == -1] = 0                               # RemoveLabelTransform(-1, 0)
    targets = []
    for st in ds_strides:                                # DownsampleSegForDSTransform2 -> resize_segmentation(order 0)
        if all(s == 1 for s in st):
            targets.append(out_s.copy())
        else:
            t = np.stack([ndi.zoom(out_s[j, 0], [1.0 / s for s in st], order=0, mode="nearest", grid_mode=True)[None] for j in range(B)])
            targets.append(t.astype(np.float32))
    return out_d, targets, margins
