"""GPU parity of the patch pipeline (csrc/augment.cu behind b2_aug_*, b200unet/augment.py) against the
CPU restatement that calls scipy.ndimage like batchgenerators does (oracle/augment.py), on the SAME plan of random parameters.

Tolerances: the kernels interpolate in fp32 (scipy in float64): data within 2e-3 of the value range; the segmentation and the
deep-supervision targets bit-exact wherever the sampling point is not within rounding of a decision boundary (a 0.5 indicator
sum or the crop border) -- at least 99.9 % of the voxels and 100 % for plans without a spatial transform."""
import numpy as np
import pytest
import torch

import util  # noqa: F401
from test_augment_host import _cases

pytestmark = pytest.mark.gpu

PATCH, STRIDES = (16, 32, 32), [(1, 1, 1), (2, 2, 2), (4, 4, 4), (4, 8, 8)]
ALL_ON = dict(p_rot=1.0, p_scale=1.0, p_noise=1.0, p_blur=1.0, p_blur_per_channel=1.0, p_brightness=1.0, p_contrast=1.0,
              p_lowres=1.0, p_lowres_per_channel=1.0, p_gamma_inverted=1.0, p_gamma=1.0)


def _compare(pipe, cases, plan, exact_seg):
    from oracle import augment as oaug
    out = pipe.run_plan(plan)
    torch.cuda.synchronize()
    od, ot, margin = oaug.apply_plan([c["data"] for c in cases], plan, PATCH, pipe.gen_patch, STRIDES)
    cd = out["data"].cpu().numpy()
    assert cd.shape == od.shape and out["keys"] == [cases[i]["key"] for i in plan["cases"]]
    scale = float(np.abs(od).max())
    safe = np.broadcast_to((np.abs(margin) > 0.05)[:, None], od.shape)       # sampling points clear of the crop border
    err = np.abs(cd - od)
    frac_bad = float((err[safe] > 2e-3 * scale).mean())
    assert frac_bad <= (0 if exact_seg else 2e-3), (frac_bad, float(err[safe].max()), scale)
    for k, (ct, ott) in enumerate(zip(out["target"], ot)):
        ct = ct.cpu().numpy()
        assert ct.shape == ott.shape, (k, ct.shape, ott.shape)
        agree = float((ct == ott).mean())
        assert agree == 1.0 if exact_seg else agree > 0.999, (k, agree)
        assert set(np.unique(ct)).issubset({0.0, 1.0, 2.0})
    # the coarser targets are exact sub-lattices of the pipeline's own full-resolution target (order-0 resize: q -> s q + s / 2)
    full = out["target"][0]
    for st, t in zip(STRIDES[1:], out["target"][1:]):
        assert torch.equal(t, full[:, :, st[0] // 2::st[0], st[1] // 2::st[1], st[2] // 2::st[2]])
    return out


@pytest.mark.parametrize("seed", [0, 1, 2])
def test_every_transform_on_matches_scipy_oracle(seed):
    from b200unet import augment
    cases = _cases(seed)
    pipe = augment.GPUPatchPipeline(cases, PATCH, 4, STRIDES, params=ALL_ON, seed=seed)
    for _ in range(2):
        plan = pipe.draw_plan()
        assert all(s["angles"] is not None and s["scale"] is not None for s in plan["spatial"])
        _compare(pipe, cases, plan, exact_seg=False)


def test_default_probabilities_many_batches():
    from b200unet import augment
    cases = _cases(5)
    pipe = augment.GPUPatchPipeline(cases, PATCH, 4, STRIDES, seed=11)
    seen = set()
    for _ in range(12):
        plan = pipe.draw_plan()
        for k in ("noise", "blur", "brightness", "contrast", "lowres", "gamma_inv", "gamma"):
            if any(v is not None for row in plan[k] for v in (row if isinstance(row, list) else [row])):
                seen.add(k)
        _compare(pipe, cases, plan, exact_seg=False)
    assert len(seen) >= 4


def test_no_augmentation_generator_is_exact_and_iterable():
    """the validation generator (no transform): bit-exact crops, labels and targets; batches come out of next()"""
    from b200unet import augment
    cases = _cases(7)
    tr, val = augment.get_moreDA_augmentation(cases, cases[:2], PATCH, deep_supervision_scales=[[1, 1, 1], [0.5, 0.5, 0.5], [0.25, 0.25, 0.25], [0.25, 0.125, 0.125]],
                                              batch_size=3, seed=4)
    plan = val.draw_plan()
    out = _compare(val, cases[:2], plan, exact_seg=True)
    assert out["data"].shape == (3, 2) + PATCH and [tuple(t.shape[2:]) for t in out["target"]] == [(16, 32, 32), (8, 16, 16), (4, 8, 8), (4, 4, 4)]
    b = next(tr)
    assert b["data"].is_cuda and torch.isfinite(b["data"]).all() and len(b["target"]) == 4 and 4 <= tr.launches_last <= 30


def test_low_resolution_simulation_alone_matches_scipy_zoom():
    """SimulateLowResolutionTransform in isolation (no other transform): nearest down / cubic up == scipy.ndimage.zoom pair;
    a voxel may pick the neighbouring source sample when (q + 0.5) p / t lands within rounding of an integer"""
    from b200unet import augment
    cases = _cases(13)
    only = dict(p_rot=0.0, p_scale=0.0, p_noise=0.0, p_blur=0.0, p_brightness=0.0, p_contrast=0.0, p_gamma_inverted=0.0, p_gamma=0.0,
                p_lowres=1.0, p_lowres_per_channel=1.0, do_mirror=False)
    pipe = augment.GPUPatchPipeline(cases, PATCH, 3, STRIDES, params=only, seed=6, prefetch=False)
    for _ in range(3):
        plan = pipe.draw_plan()
        assert all(z is not None and 0.5 <= z <= 1.0 for row in plan["lowres"] for z in row)
        _compare(pipe, cases, plan, exact_seg=False)
    plan["lowres"] = [[0.5, 1.0]] * 3                    # zoom 1: the spline interpolates its own samples;  zoom 0.5: every second voxel
    out = _compare(pipe, cases, plan, exact_seg=False)
    plain = pipe.run_plan(dict(plan, lowres=[[None, None]] * 3))
    assert torch.allclose(out["data"][:, 1], plain["data"][:, 1], atol=1e-4)
    assert float((out["data"][:, 0] - plain["data"][:, 0]).abs().max()) > 0.1


def test_retain_stats_gamma_keeps_mean_and_std():
    from b200unet import augment
    cases = _cases(9)
    off = dict(p_rot=0.0, p_scale=0.0, p_noise=0.0, p_blur=0.0, p_brightness=0.0, p_contrast=0.0, p_gamma_inverted=1.0, p_gamma=1.0,
               do_mirror=False)
    pipe = augment.GPUPatchPipeline(cases, PATCH, 2, [(1, 1, 1)], params=off, seed=2)
    plan = pipe.draw_plan()
    plain = dict(plan, gamma=[[None, None]] * 2, gamma_inv=[[None, None]] * 2)
    a, b = pipe.run_plan(plain)["data"], pipe.run_plan(plan)["data"]
    assert not torch.allclose(a, b)
    assert torch.allclose(a.mean((2, 3, 4)), b.mean((2, 3, 4)), atol=1e-4) and torch.allclose(a.std((2, 3, 4)), b.std((2, 3, 4)), rtol=1e-3)


def test_trainer_consumes_the_pipeline():
    """the generators feed run_iteration directly (device-resident batches)"""
    from b200unet import augment
    from b200unet.configs import CONFIGS
    from b200unet.trainers import nnUNetTrainerSequential
    geom = CONFIGS["tiny"]
    cases = [{"key": c["key"], "data": c["data"][[0, 2]]} for c in _cases(3)]           # one input channel + segmentation
    strides, cum = [(1, 1, 1)], [1, 1, 1]
    for k in geom.pool[:-1]:
        cum = [a * b for a, b in zip(cum, k)]
        strides.append(tuple(cum))
    pipe = augment.GPUPatchPipeline(cases, geom.patch, geom.batch, strides, seed=1)
    tr = nnUNetTrainerSequential(geom, precision="fp32", task="A")
    tr.initialize()
    losses = [float(tr.run_iteration(pipe)) for _ in range(4)]
    assert all(np.isfinite(l) for l in losses)


def test_rehearsal_trainer_trains_from_the_fused_pipeline():
    """reference rehearsal:65-173 with the GPU pipeline in the place of DataLoader3D: fused case list -> batches -> iterations"""
    from b200unet.configs import CONFIGS
    from b200unet.trainers import nnUNetTrainerRehearsal
    geom = CONFIGS["tiny"]
    mk = lambda seed, tag: {"%s%d" % (tag, i): {"data": c["data"][[0, 2]]} for i, c in enumerate(_cases(seed))}
    tr = nnUNetTrainerRehearsal(geom, precision="fp32", task="A", samples_in_perc=0.67)
    tr.initialize()
    tr.start_task("B")
    gen_tr, gen_val = tr.get_basic_generators(mk(1, "b"), {"A": mk(2, "a")}, mk(3, "v"), seed=5)
    assert sorted(gen_tr.keys) == ["a%d" % i for i in sorted(int(k[1]) for k in gen_tr.keys if k[0] == "a")] + ["b0", "b1", "b2"]
    assert sum(k[0] == "a" for k in gen_tr.keys) == 2 and gen_val.keys == ["v0", "v1", "v2"] and not gen_val.train
    seen = set()
    for _ in range(6):
        b = next(gen_tr)
        seen.update(b["keys"])
    assert any(k.startswith("a") for k in seen) and any(k.startswith("b") for k in seen)
    assert all(np.isfinite(float(tr.run_iteration(gen_tr))) for _ in range(3))
    assert np.isfinite(float(tr.run_iteration(gen_val, do_backprop=False)))


def test_pipeline_matches_the_committed_fixture():
    """the CUDA pipeline on the plan stored in tests/golden/augment_tiny.npz (every transform on) against the oracle output
    committed beside it -- same tolerances as the live comparison above"""
    from b200unet import augment
    cases, plan, patch, strides, gen_patch, params, data, targets, margin = util.load_augment_fixture()
    pipe = augment.GPUPatchPipeline(cases, patch, len(plan["cases"]), strides, params=params, seed=0, prefetch=False)
    assert pipe.gen_patch == gen_patch
    out = pipe.run_plan(plan)
    torch.cuda.synchronize()
    bad, agree = util.augment_mismatch(out["data"].cpu().numpy(), [t.cpu().numpy() for t in out["target"]], data, targets, margin)
    assert bad < 2e-3 and agree > 0.999, (bad, agree)
