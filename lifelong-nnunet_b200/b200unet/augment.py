"""GPU-side patch pipeline -- the generator pair the reference obtains from nnunet's ``DataLoader3D`` +
batchgenerators' ``get_moreDA_augmentation`` (call sites training/network_training/multihead/nnUNetTrainerMultiHead.py:505-511
and :904-922; both packages are un-vendored: semantics restated from nnunet@77bc485 / batchgenerators 0.21 / scipy.ndimage).

At 300+ patches/s per GPU the CPU augmenter (~12 worker processes, tens of patches/s) is an order of magnitude too slow, so the
preprocessed cases live in HBM ([C + 1, D, H, W] fp32, segmentation as the last channel -- nnunet's ``.npy`` layout) and one
batch is a handful of kernel launches behind the C ABI (csrc/augment.cu, ``b2_aug_*``):

  crop with foreground oversampling (DataLoader3D) -> SpatialTransform (rotation +-30 deg / scaling 0.7-1.4, p 0.2 each;
  data order 3, segmentation order 1 per label) -> GaussianNoise (p 0.1) -> GaussianBlur (p 0.2, per channel 0.5) ->
  BrightnessMultiplicative (p 0.15) -> ContrastAugmentation (p 0.15) -> SimulateLowResolution (p 0.25, per channel 0.5: nearest
  down to 50-100 %, cubic up) -> Gamma on the inverted image (p 0.1) -> Gamma (p 0.3), both with retain_stats -> Mirror -> RemoveLabel(-1, 0) -> deep-supervision targets (order 0) -> {'data', 'target', 'keys'}

The HOST draws every random number (`draw_plan`, numpy RandomState, order documented there) into a small plain-dict "plan";
`run_plan` turns a plan into kernel launches.  tests/ hands the same plan to the CPU restatement (oracle/augment.py, which calls
scipy.ndimage like batchgenerators does) and compares.  Not implemented: elastic deformation (off in nnU-Net's 3D default), the
cascade / mask transforms.
"""
import ctypes as C
import math

import numpy as np
import torch

from . import _lib

TWO_PI = 2.0 * np.pi
DEFAULT_3D_PARAMS = {           # nnunet default_3D_augmentation_params as set up by nnUNetTrainerV2.setup_DA_params
    "rotation_x": (-30. / 360 * TWO_PI, 30. / 360 * TWO_PI), "rotation_y": (-30. / 360 * TWO_PI, 30. / 360 * TWO_PI),
    "rotation_z": (-30. / 360 * TWO_PI, 30. / 360 * TWO_PI), "p_rot": 0.2,
    "scale_range": (0.7, 1.4), "p_scale": 0.2,
    "p_noise": 0.1, "noise_variance": (0.0, 0.1),
    "p_blur": 0.2, "blur_sigma": (0.5, 1.0), "p_blur_per_channel": 0.5,
    "p_brightness": 0.15, "brightness_range": (0.75, 1.25),
    "p_contrast": 0.15, "contrast_range": (0.75, 1.25),
    "p_lowres": 0.25, "lowres_zoom": (0.5, 1.0), "p_lowres_per_channel": 0.5,
    "p_gamma_inverted": 0.1, "p_gamma": 0.3, "gamma_range": (0.7, 1.5),
    "do_mirror": True, "mirror_axes": (0, 1, 2),
}


def rotation_matrix(ax, ay, az):
    """batchgenerators rotate_coords_3d: coords_row . (Rx . Ry . Rz)"""
    cx, sx, cy, sy, cz, sz = math.cos(ax), math.sin(ax), math.cos(ay), math.sin(ay), math.cos(az), math.sin(az)
    rx = np.array([[1, 0, 0], [0, cx, -sx], [0, sx, cx]], dtype=np.float64)
    ry = np.array([[cy, 0, sy], [0, 1, 0], [-sy, 0, cy]], dtype=np.float64)
    rz = np.array([[cz, -sz, 0], [sz, cz, 0], [0, 0, 1]], dtype=np.float64)
    return rx @ ry @ rz


def get_patch_size(final_patch_size, rot_x, rot_y, rot_z, scale_range):
    """nnunet default_data_augmentation.get_patch_size: the generator patch that still covers the final patch after the largest
    rotation / smallest zoom"""
    lim = lambda r: min(90. / 360 * TWO_PI, max(abs(r[0]), abs(r[1])) if isinstance(r, (tuple, list)) else r)
    rot_x, rot_y, rot_z = lim(rot_x), lim(rot_y), lim(rot_z)
    coords = np.array(final_patch_size, dtype=np.float64)
    final = coords.copy()
    for a in ((rot_x, 0, 0), (0, rot_y, 0), (0, 0, rot_z)):
        final = np.max(np.vstack((np.abs(coords @ rotation_matrix(*a)), final)), 0)
    final /= min(scale_range)
    return tuple(int(v) for v in final.astype(int))


def class_locations_of(seg, num_samples=10000, seed=1234):
    """nnunet preprocessing: up to 10000 random voxel coordinates per foreground class (GenericPreprocessor._run_internal)"""
    rs = np.random.RandomState(seed)
    out = {}
    for c in np.unique(seg):
        if c <= 0:
            continue
        loc = np.argwhere(seg == c)
        n = min(num_samples, len(loc))
        n = max(n, int(np.ceil(len(loc) * 0.01)))
        out[int(c)] = loc[rs.choice(len(loc), n, replace=False)]
    return out


def _two_sided(rs, rng):
    """batchgenerators' draw for scale / contrast / gamma: with p 0.5 (and a range that reaches below 1) from [lo, 1), else
    from [max(lo, 1), hi)"""
    if rs.random_sample() < 0.5 and rng[0] < 1:
        return float(rs.uniform(rng[0], 1))
    return float(rs.uniform(max(rng[0], 1), rng[1]))


class GPUPatchPipeline:
    """Iterator of training (or validation) batches.  `cases`: list of dicts {'key', 'data': ndarray [C + 1, D, H, W] fp32 with the
    segmentation as last channel, optional 'class_locations': {label: int array [N, 3]}}."""

    def __init__(self, cases, patch_size, batch_size, ds_strides, params=None, oversample_foreground_percent=0.33, seed=1234,
                 device=None, train=True, plan_only=False, prefetch=True, prefetch_delay=0.0):
        """plan_only: host-side use (draw_plan) without a device -- run_plan then raises.
        prefetch: next() returns the batch produced during the PREVIOUS call and enqueues the following one on the pipeline's own
        stream, so augmentation overlaps the consumer's training step (same plans, same order as without prefetch)."""
        self.plan_only = bool(plan_only)
        self.prefetch, self._pending, self._stream = bool(prefetch), None, None
        # optional: the producer thread starts its (GIL-holding) parameter draws this long after next() returned (measured: no
        # gain at cfg2 -- the cost of feeding a training step from the pipeline is the GPU time of its kernels, not host contention)
        self.prefetch_delay = float(prefetch_delay)
        if not self.plan_only:
            self.lib = _lib.load()
            self.device = torch.device("cuda", torch.cuda.current_device()) if device is None else torch.device(device)
            if self.device.type != "cuda":
                raise RuntimeError("GPUPatchPipeline runs on a CUDA device only (sm_100a); there is no CPU fallback")
        self.params = dict(DEFAULT_3D_PARAMS)
        self.params.update(params or {})
        self.patch = tuple(int(p) for p in patch_size)
        self.train = bool(train)
        p = self.params
        self.gen_patch = get_patch_size(self.patch, p["rotation_x"], p["rotation_y"], p["rotation_z"], (0.85, 1.25)) \
            if self.train else self.patch
        self.batch_size = int(batch_size)
        self.ds_strides = [tuple(int(s) for s in st) for st in ds_strides]
        self.oversample = float(oversample_foreground_percent)
        self.rs = np.random.RandomState(seed)
        self.keys, self.vols, self.shapes, self.locs = [], [], [], []
        for c in cases:
            d = np.ascontiguousarray(c["data"], dtype=np.float32)
            assert d.ndim == 4, "case data must be [C + 1, D, H, W] with the segmentation as last channel"
            self.keys.append(c["key"])
            self.shapes.append(tuple(int(s) for s in d.shape[1:]))
            self.locs.append(c.get("class_locations") or class_locations_of(d[-1]))
            self.vols.append(None if self.plan_only else torch.from_numpy(d).to(self.device))
        self.C = int(np.shape(cases[0]["data"])[0]) - 1
        if self.batch_size > _lib.AUG_MAX_SAMPLES or self.batch_size * self.C > _lib.AUG_MAX_BC or len(self.ds_strides) > _lib.AUG_MAX_SCALES:
            raise NotImplementedError("b2_aug_*: at most %d samples, %d (sample, channel) pairs and %d deep-supervision scales per batch"
                                      % (_lib.AUG_MAX_SAMPLES, _lib.AUG_MAX_BC, _lib.AUG_MAX_SCALES))
        if self.plan_only:
            return
        B, Cc, g, pp = self.batch_size, self.C, self.gen_patch, self.patch
        f32 = dict(dtype=torch.float32, device=self.device)
        self._crop_data, self._crop_seg = torch.empty((B, Cc) + g, **f32), torch.empty((B, 1) + g, **f32)
        self._data, self._tmp, self._seg = torch.empty((B, Cc) + pp, **f32), torch.empty((B, Cc) + pp, **f32), torch.empty((B, 1) + pp, **f32)
        V = pp[0] * pp[1] * pp[2]
        self._stats_a, self._stats_b = torch.empty((B * Cc, 4), **f32), torch.empty((B * Cc, 4), **f32)
        self._scr = torch.empty(int(self.lib.b2_aug_stats_scratch_bytes(B * Cc, V)), dtype=torch.uint8, device=self.device)
        self._lowres_scr = torch.empty(int(self.lib.b2_aug_lowres_scratch_bytes(C.byref((C.c_int32 * 3)(*pp)))), dtype=torch.uint8,
                                       device=self.device)
        self.launches_last = 0

    # -- host: every random number of one batch, in this order -------------------------------------------------------------
    def draw_plan(self):
        rs, p, B, Cc = self.rs, self.params, self.batch_size, self.C
        gen, patch = np.array(self.gen_patch), np.array(self.patch)
        plan = {"cases": [int(i) for i in rs.choice(len(self.keys), B, True)], "lb": [], "spatial": [], "noise": [], "blur": [],
                "brightness": [], "contrast": [], "lowres": [], "gamma_inv": [], "gamma": [], "flips": [],
                "seed": int(rs.randint(0, 2 ** 31 - 1))}
        for j, ci in enumerate(plan["cases"]):
            # DataLoader3D.generate_train_batch: the last round(B * oversample) samples are forced to contain foreground
            force_fg = not (j < round(B * (1 - self.oversample)))
            shape = np.array(self.shapes[ci])
            need = gen - patch
            for d in range(3):
                if need[d] + shape[d] < gen[d]:
                    need[d] = gen[d] - shape[d]
            lb = -need // 2
            ub = shape + need // 2 + need % 2 - gen
            locs = {c: v for c, v in self.locs[ci].items() if len(v)}
            if force_fg and locs:
                cls = sorted(locs)[int(rs.choice(len(locs)))]
                vox = locs[cls][int(rs.choice(len(locs[cls])))]
                box = [int(max(lb[d], vox[d] - gen[d] // 2)) for d in range(3)]
            else:
                box = [int(rs.randint(lb[d], ub[d] + 1)) for d in range(3)]
            plan["lb"].append(box)
        for j in range(B):
            sp = {"angles": None, "scale": None}
            if self.train:
                if rs.uniform() < p["p_rot"]:
                    sp["angles"] = [float(rs.uniform(*p["rotation_x"])), float(rs.uniform(*p["rotation_y"])), float(rs.uniform(*p["rotation_z"]))]
                if rs.uniform() < p["p_scale"]:
                    sp["scale"] = _two_sided(rs, p["scale_range"])
            plan["spatial"].append(sp)
        for j in range(B):
            on = self.train
            plan["noise"].append(float(rs.uniform(*p["noise_variance"])) if on and rs.uniform() < p["p_noise"] else None)
            plan["blur"].append([float(rs.uniform(*p["blur_sigma"])) if rs.uniform() <= p["p_blur_per_channel"] else None for _ in range(Cc)]
                                if on and rs.uniform() < p["p_blur"] else [None] * Cc)
            plan["brightness"].append([float(rs.uniform(*p["brightness_range"])) for _ in range(Cc)]
                                      if on and rs.uniform() < p["p_brightness"] else [None] * Cc)
            plan["contrast"].append([_two_sided(rs, p["contrast_range"]) for _ in range(Cc)]
                                    if on and rs.uniform() < p["p_contrast"] else [None] * Cc)
            plan["lowres"].append([float(rs.uniform(*p["lowres_zoom"])) if rs.uniform() < p["p_lowres_per_channel"] else None
                                   for _ in range(Cc)] if on and rs.uniform() < p["p_lowres"] else [None] * Cc)
            plan["gamma_inv"].append([_two_sided(rs, p["gamma_range"]) for _ in range(Cc)]
                                     if on and rs.uniform() < p["p_gamma_inverted"] else [None] * Cc)
            plan["gamma"].append([_two_sided(rs, p["gamma_range"]) for _ in range(Cc)]
                                 if on and rs.uniform() < p["p_gamma"] else [None] * Cc)
            fl = 0
            if on and p["do_mirror"]:
                for a in (0, 1, 2):
                    if a in p["mirror_axes"] and rs.uniform() < 0.5:
                        fl |= 1 << a
            plan["flips"].append(fl)
        return plan

    # -- device ----------------------------------------------------------------------------------------------------------------
    def run_plan(self, plan):
        if self.plan_only:
            raise RuntimeError("GPUPatchPipeline(plan_only=True) holds no device state; there is no CPU fallback")
        lib, B, Cc, dev = self.lib, self.batch_size, self.C, self.device
        st = C.c_void_p(torch.cuda.current_stream(dev).cuda_stream)
        n0 = _lib.launch_count()
        g3, p3 = (C.c_int32 * 3)(*self.gen_patch), (C.c_int32 * 3)(*self.patch)
        V = self.patch[0] * self.patch[1] * self.patch[2]
        ptr = lambda t: C.c_void_p(t.data_ptr())
        cases = (_lib.AugCase * B)()
        for j, ci in enumerate(plan["cases"]):
            cases[j].volume = self.vols[ci].data_ptr()
            cases[j].dhw = (C.c_int32 * 3)(*self.shapes[ci])
            cases[j].lb = (C.c_int32 * 3)(*plan["lb"][j])
            sp = plan["spatial"][j]
            if sp["angles"] is None and sp["scale"] is None:      # only the centre window is ever read
                lo = [(g - q) // 2 for g, q in zip(self.gen_patch, self.patch)]
                cases[j].win_lo, cases[j].win_hi = (C.c_int32 * 3)(*lo), (C.c_int32 * 3)(*[l + q for l, q in zip(lo, self.patch)])
            else:
                cases[j].win_lo, cases[j].win_hi = (C.c_int32 * 3)(0, 0, 0), (C.c_int32 * 3)(*self.gen_patch)
        _lib.check(lib.b2_aug_crop(cases, B, Cc, C.byref(g3), ptr(self._crop_data), ptr(self._crop_seg), st))
        tf = (_lib.AugSpatial * B)()
        for j, sp in enumerate(plan["spatial"]):
            tf[j].modified = int(sp["angles"] is not None or sp["scale"] is not None)
            m = np.eye(3)
            if sp["angles"] is not None:
                m = rotation_matrix(*sp["angles"]).T         # column form of coords_row . R
            if sp["scale"] is not None:
                m = m * sp["scale"]
            tf[j].m = (C.c_float * 9)(*[float(v) for v in m.reshape(-1)])
            tf[j].ctr = (C.c_float * 3)(*[g / 2. - 0.5 for g in self.gen_patch])
            tf[j].lb = (C.c_int32 * 3)(*[(g - q) // 2 for g, q in zip(self.gen_patch, self.patch)])
        _lib.check(lib.b2_aug_spatial(tf, B, Cc, C.byref(g3), C.byref(p3), ptr(self._crop_data), ptr(self._crop_seg), ptr(self._data),
                                      ptr(self._seg), st))

        def pointwise(op, values, p1=0.0, stats=False):
            if not any(v is not None for row in values for v in row):
                return
            ops = (_lib.AugOp * (B * Cc))()
            for j in range(B):
                for c in range(Cc):
                    v = values[j][c]
                    ops[j * Cc + c].op = _lib.AUG_NONE if v is None else op
                    ops[j * Cc + c].p = (C.c_float * 3)(0.0 if v is None else float(v), p1, 0.0)
            sa = ptr(self._stats_a) if stats else None
            if stats:
                _lib.check(lib.b2_aug_stats(ptr(self._data), B * Cc, V, sa, ptr(self._scr), st))
            _lib.check(lib.b2_aug_pointwise(ops, B * Cc, V, ptr(self._data), sa, None, C.c_uint64(plan["seed"]), st))
            if op == _lib.AUG_GAMMA_A:          # retain_stats: statistics after the curve, then the rescale
                for i in range(B * Cc):
                    if ops[i].op == _lib.AUG_GAMMA_A:
                        ops[i].op = _lib.AUG_GAMMA_B
                _lib.check(lib.b2_aug_stats(ptr(self._data), B * Cc, V, ptr(self._stats_b), ptr(self._scr), st))
                _lib.check(lib.b2_aug_pointwise(ops, B * Cc, V, ptr(self._data), sa, ptr(self._stats_b), C.c_uint64(plan["seed"]), st))

        pointwise(_lib.AUG_NOISE, [[s] * Cc for s in plan["noise"]])
        if any(s is not None for row in plan["blur"] for s in row):
            taps = (_lib.AugBlur * (B * Cc))()
            for j in range(B):
                for c in range(Cc):
                    s = plan["blur"][j][c]
                    if s is None:
                        continue
                    r = int(4.0 * s + 0.5)
                    w = np.exp(-0.5 / (s * s) * np.arange(-r, r + 1, dtype=np.float64) ** 2)
                    w /= w.sum()
                    taps[j * Cc + c].radius = r
                    taps[j * Cc + c].w = (C.c_float * 5)(*([float(v) for v in w[r:]] + [0.0] * (4 - r)))
            _lib.check(lib.b2_aug_blur(taps, B * Cc, C.byref(p3), ptr(self._data), ptr(self._tmp), st))
        pointwise(_lib.AUG_MUL, plan["brightness"])
        pointwise(_lib.AUG_CONTRAST, plan["contrast"], stats=True)
        for j in range(B):                       # SimulateLowResolutionTransform: one (sample, channel) volume per call
            for c in range(Cc):
                zoom = plan.get("lowres", [[None] * Cc] * B)[j][c]
                if zoom is None:
                    continue
                t3 = (C.c_int32 * 3)(*[max(2, int(np.round(q * zoom))) for q in self.patch])
                _lib.check(lib.b2_aug_lowres(ptr(self._data[j, c]), C.byref(p3), C.byref(t3), ptr(self._lowres_scr), st))
        pointwise(_lib.AUG_GAMMA_A, plan["gamma_inv"], p1=1.0, stats=True)
        pointwise(_lib.AUG_GAMMA_A, plan["gamma"], p1=0.0, stats=True)
        data = torch.empty_like(self._data)
        targets = [torch.empty((B, 1) + tuple(q // s for q, s in zip(self.patch, stv)), dtype=torch.float32, device=dev)
                   for stv in self.ds_strides]
        tp = (C.c_void_p * len(targets))(*[t.data_ptr() for t in targets])
        strides = (C.c_int32 * (3 * len(targets)))(*[s for stv in self.ds_strides for s in stv])
        flips = (C.c_int32 * B)(*plan["flips"])
        _lib.check(lib.b2_aug_finalize(flips, B, Cc, C.byref(p3), ptr(self._data), ptr(self._seg), ptr(data), tp, strides, len(targets), st))
        self.launches_last = _lib.launch_count() - n0
        return {"data": data, "target": targets, "keys": [self.keys[i] for i in plan["cases"]]}

    def __iter__(self):
        return self

    def _produce(self):
        """one batch on the pipeline's own stream; returns (batch, event)"""
        dev = self.device
        torch.cuda.set_device(dev)
        if self._stream is None:
            self._stream = torch.cuda.Stream(dev)
        with torch.cuda.stream(self._stream):
            batch = self.run_plan(self.draw_plan())
            ev = torch.cuda.Event()
            ev.record(self._stream)
        return batch, ev

    def _produce_async(self):
        """draw + enqueue the next batch on a host thread, so neither the parameter draws nor the launches delay the consumer's
        own enqueue work (ctypes calls and stream synchronisation release the GIL)"""
        import threading
        box = {}

        def work():
            try:
                if self.prefetch_delay > 0:
                    import time
                    time.sleep(self.prefetch_delay)
                box["out"] = self._produce()
            except BaseException as e:      # re-raised by the consumer in __next__
                box["err"] = e
        t = threading.Thread(target=work, daemon=True)
        t.start()
        return t, box

    def __next__(self):
        if not self.prefetch:
            return self.run_plan(self.draw_plan())
        if self._pending is None:
            batch, ev = self._produce()
        else:
            t, box = self._pending
            t.join()
            if "err" in box:
                self._pending = None
                raise box["err"]
            batch, ev = box["out"]
        cur = torch.cuda.current_stream(self.device)
        cur.wait_event(ev)
        for t_ in [batch["data"]] + batch["target"]:
            t_.record_stream(cur)
        self._pending = self._produce_async()
        return batch


def get_moreDA_augmentation(cases_train, cases_val, patch_size, params=None, deep_supervision_scales=None, batch_size=2,
                            oversample_foreground_percent=0.33, seed=1234, device=None):
    """Signature-level stand-in for the generator pair of nnunet's get_moreDA_augmentation(dl_tr, dl_val, patch_size, params,
    deep_supervision_scales=...): (training generator with augmentation, validation generator without).  deep_supervision_scales
    are nnU-Net's fractions ([[1, 1, 1], [0.5, 0.5, 0.5], ...]); they become integer strides here."""
    scales = deep_supervision_scales or [[1, 1, 1]]
    strides = [tuple(int(round(1.0 / s)) for s in sc) for sc in scales]
    tr = GPUPatchPipeline(cases_train, patch_size, batch_size, strides, params, oversample_foreground_percent, seed, device, train=True)
    val = GPUPatchPipeline(cases_val, patch_size, batch_size, strides, params, oversample_foreground_percent, seed + 1, device, train=False)
    return tr, val
