"""CPU: host logic of the GPU patch pipeline (b200unet/augment.py: generator patch size, foreground oversampling, parameter
draws) and the oracle's own building blocks (oracle/augment.py) against scipy / closed forms."""
import numpy as np
import pytest

import util  # noqa: F401
from b200unet import augment
from oracle import augment as oaug


def _cases(seed=0, shapes=((40, 60, 56), (24, 70, 64), (50, 48, 80)), C=2):
    rs = np.random.RandomState(seed)
    out = []
    for i, sh in enumerate(shapes):
        d = rs.randn(C + 1, *sh).astype(np.float32)
        seg = np.zeros(sh, np.float32)
        seg[sh[0] // 4: sh[0] // 2, sh[1] // 3: sh[1] // 2, sh[2] // 4: sh[2] // 2] = 1
        seg[sh[0] // 2: sh[0] // 2 + 4, 5:20, 8:30] = 2
        seg[:2] = -1                                   # outside the nonzero mask (nnunet marks it -1)
        d[-1] = seg
        out.append({"key": "case_%d" % i, "data": d})
    return out


def test_generator_patch_size_known_answer():
    """nnunet get_patch_size for the cfg2 patch, +-30 degrees on every axis, scale range (0.85, 1.25): (64, 128, 128) grows to the
    box that still covers the patch after the largest rotation, divided by the smallest zoom"""
    p = augment.DEFAULT_3D_PARAMS
    g = augment.get_patch_size((64, 128, 128), p["rotation_x"], p["rotation_y"], p["rotation_z"], (0.85, 1.25))
    c, s = np.cos(np.pi / 6), np.sin(np.pi / 6)
    expect = np.array([64 * c + 128 * s, 128 * c + 128 * s, 64 * s + 128 * c]) / 0.85
    assert g == tuple(int(v) for v in expect)
    assert np.allclose(augment.rotation_matrix(0.3, -0.2, 0.5), oaug.rotation_matrix(0.3, -0.2, 0.5))


def test_plan_draws_follow_the_dataloader_rules():
    cases = _cases()
    pipe = augment.GPUPatchPipeline(cases, (16, 32, 32), 4, [(1, 1, 1), (2, 2, 2)], seed=3, plan_only=True)
    gen = np.array(pipe.gen_patch)
    n_fg = n_mod = n_flip = 0
    for _ in range(200):
        plan = pipe.draw_plan()
        for j, (ci, lb) in enumerate(zip(plan["cases"], plan["lb"])):
            shape = np.array(cases[ci]["data"].shape[1:])
            need = np.maximum(gen - np.array(pipe.patch), gen - shape)
            assert all(lb[d] >= -need[d] // 2 - 0 for d in range(3))
            if j >= round(4 * (1 - 0.33)):             # forced-foreground samples: the box holds a foreground voxel near its centre
                box = cases[ci]["data"][-1][tuple(slice(max(0, l), max(0, l + g)) for l, g in zip(lb, gen))]
                n_fg += int((box > 0).any())
        n_mod += sum(int(s["angles"] is not None or s["scale"] is not None) for s in plan["spatial"])
        n_flip += sum(bin(f).count("1") for f in plan["flips"])
        for s in plan["spatial"]:
            if s["scale"] is not None:
                assert 0.7 <= s["scale"] <= 1.4
            if s["angles"] is not None:
                assert all(abs(a) <= np.pi / 6 + 1e-9 for a in s["angles"])
    assert n_fg == 200                                 # one forced sample per batch of 4 (round(4 * 0.67) = 3)
    assert 0.25 < n_mod / 800 < 0.47                   # 1 - 0.8^2 = 0.36
    assert 0.45 < n_flip / 2400 < 0.55
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        pipe.run_plan(plan)
    val = augment.GPUPatchPipeline(cases, (16, 32, 32), 2, [(1, 1, 1)], seed=3, train=False, plan_only=True)
    vp = val.draw_plan()
    assert val.gen_patch == (16, 32, 32) and all(s == {"angles": None, "scale": None} for s in vp["spatial"]) and vp["flips"] == [0, 0]


def test_oracle_crop_pads_with_constants():
    cases = _cases()
    d, s = oaug.crop_case(cases[1]["data"], [-3, 60, -2], (30, 20, 70))
    assert d.shape == (2, 30, 20, 70) and s.shape == (1, 30, 20, 70)
    assert (d[:, :3] == 0).all() and (s[:, :3] == -1).all() and (s[:, :, 10:] == -1).all() and (d[:, :, 10:] == 0).all()
    assert np.array_equal(d[:, 3:27, :10, 2:66], cases[1]["data"][:2, :24, 60:70, :64])


def test_oracle_noise_generator_is_standard_normal_and_counter_based():
    a = oaug.normal_field(77, 0, 200000)
    assert abs(a.mean()) < 0.01 and abs(a.std() - 1) < 0.01
    assert np.array_equal(oaug.normal_field(77, 1000, 50), a[1000:1050])
    assert not np.array_equal(oaug.normal_field(78, 0, 50), a[:50])


def test_oracle_identity_plan_is_a_centre_crop_and_ds_targets_pick_odd_voxels():
    cases = _cases()
    plan = {"cases": [0], "lb": [[2, 3, 4]], "spatial": [{"angles": None, "scale": None}], "noise": [None], "blur": [[None, None]],
            "brightness": [[None, None]], "contrast": [[None, None]], "lowres": [[None, None]], "gamma_inv": [[None, None]],
            "gamma": [[None, None]],
            "flips": [0], "seed": 1}
    d, t, _ = oaug.apply_plan([c["data"] for c in cases], plan, (16, 32, 32), (20, 40, 36), [(1, 1, 1), (2, 2, 2), (4, 4, 4)])
    src = cases[0]["data"]
    assert np.array_equal(d[0], src[:2, 4:20, 7:39, 6:38])
    seg = src[2, 4:20, 7:39, 6:38].copy()
    seg[seg == -1] = 0
    assert np.array_equal(t[0][0, 0], seg)
    # scipy zoom(order 0, grid_mode) == skimage resize(order 0): output voxel q reads input 2q + 1 (resp. 4q + 2)
    assert np.array_equal(t[1][0, 0], seg[1::2, 1::2, 1::2]) and np.array_equal(t[2][0, 0], seg[2::4, 2::4, 2::4])
