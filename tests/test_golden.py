"""CPU: the oracle restatements (oracle/cl_losses.py, oracle/step.py) against the committed golden fixtures that
oracle/gen_golden.py produced from the reference's own files.  Runs everywhere (no reference, no GPU needed)."""
import os

import numpy as np
import pytest
import torch

import util
from oracle import cl_losses
from oracle.gen_golden import seeded_case

GOLD = os.path.join(util.ROOT, "tests", "golden")
RT = 2e-6


@pytest.fixture(scope="module")
def case():
    return seeded_case()


@pytest.fixture(scope="module")
def gold():
    return np.load(os.path.join(GOLD, "cl_losses.npz"))


def _base(case):
    return cl_losses.multiple_output_loss2(case["xs"], case["ys"], case["weights"])


def test_base_loss(case, gold):
    xs = [x.clone().requires_grad_() for x in case["xs"]]
    v = cl_losses.multiple_output_loss2(xs, case["ys"], case["weights"])
    v.backward()
    assert abs(v.item() - gold["base"]) < RT * abs(gold["base"])
    np.testing.assert_allclose(xs[0].grad.numpy(), gold["base_dx0"], rtol=1e-5, atol=1e-9)


def test_ewc_q1_generator_and_list(case, gold):
    f2 = {t: case["fisher"][t] for t in ("A", "B")}
    p2 = {t: case["params"][t] for t in ("A", "B")}
    for tag, strict in (("gen", True), ("list", False)):
        ps = [(n, p.clone().requires_grad_()) for n, p in case["named"]]
        v = _base(case) + cl_losses.ewc_penalty(ps, f2, p2, 0.4, strict_reference=strict)
        v.backward()
        assert abs(v.item() - gold["ewc_" + tag]) < RT * abs(gold["ewc_" + tag])
        np.testing.assert_allclose(ps[2][1].grad.numpy(), gold["ewc_%s_dp2" % tag], rtol=1e-5, atol=1e-9)
    assert gold["ewc_list"] > gold["ewc_gen"] > gold["base"]


def test_rw_q2(case, gold):
    ps = [(n, p.clone().requires_grad_()) for n, p in case["named"]]
    v = _base(case) + cl_losses.rw_penalty(ps, case["fisher"], case["params"], case["scores"], 0.4, True)
    v.backward()
    assert abs(v.item() - gold["rw_first"]) < RT * abs(gold["rw_first"])
    np.testing.assert_allclose(ps[2][1].grad.numpy(), gold["rw_dp2"], rtol=1e-5, atol=1e-9)
    assert abs(gold["rw_second"] - gold["base"]) < 1e-7      # Q2: exhausted generator -> no penalty


def test_lwf_mib(case, gold):
    v = _base(case) + cl_losses.lwf_distillation(case["xs"][0], case["xo"][0], 2.0)
    assert abs(v.item() - gold["lwf"]) < 1e-5 * abs(gold["lwf"])
    xs = [x.clone().requires_grad_() for x in case["xs"]]
    v = cl_losses.mib_loss(xs, case["xo"], case["ys"], case["weights"], 0.9, 10)
    v.backward()
    assert abs(v.item() - gold["mib"]) < 1e-5 * abs(gold["mib"])
    np.testing.assert_allclose(xs[0].grad.numpy(), gold["mib_dx0"], rtol=1e-4, atol=1e-8)


def test_pod_q5_q6(case, gold):
    for k in case["layers"]:
        v = cl_losses.local_pod(case["layers"][k], case["layers_old"][k], 3)
        assert abs(v.item() - gold["pod_" + k]) < RT * abs(gold["pod_" + k]), k
    five = {k: v for k, v in case["layers"].items() if v.dim() == 5}
    five_old = {k: case["layers_old"][k] for k in five}
    v = _base(case) + cl_losses.pod_running(five, five_old, 1e-2, 3)
    assert abs(v.item() - gold["pod_total"]) < RT * abs(gold["pod_total"])


def test_plop_pseudo_q10(case, gold):
    for i in range(3):
        x = case["xs"][i].clone().requires_grad_()
        v = cl_losses.plop_pseudo_label_loss(x, case["xo"][i], case["ys"][i].squeeze(), case["thr"][i], 1.0)
        g = gold["plop_pseudo_%d" % i]
        if np.isnan(g):
            assert torch.isnan(v)     # a (b, w) column without background voxels: 0/0 like the reference
            continue
        v.backward()
        assert abs(v.item() - g) < 1e-5 * abs(g)
        np.testing.assert_allclose(x.grad.numpy(), gold["plop_pseudo_%d_dx" % i], rtol=1e-4, atol=1e-8)


def test_unet_tiny_fixture():
    from b200unet import synth
    from b200unet.configs import CONFIGS
    from oracle import step
    g = np.load(os.path.join(GOLD, "unet_tiny.npz"))
    geom = CONFIGS["tiny"]
    net = step.build_network(geom.in_channels, geom.base_features, geom.num_classes, [list(k) for k in geom.pool])
    data, targets = synth.make_batch(geom)
    out = net(data)
    l = cl_losses.multiple_output_loss2(out, targets, cl_losses.ds_loss_weights(geom.num_pool))
    assert abs(l.item() - g["loss"]) < 1e-4 * abs(g["loss"])
    np.testing.assert_allclose(out[-1].detach().numpy(), g["logits_last"], rtol=1e-3, atol=1e-4)


def test_ds_weights_and_hard_dice():
    w = cl_losses.ds_loss_weights(5)
    assert np.allclose(w, [0.53333333, 0.26666667, 0.13333333, 0.06666667, 0.0])
    assert cl_losses.ds_loss_weights(2) == [1.0, 0.0]
    lg = torch.zeros(1, 3, 2, 2, 2)
    lg[:, 1] = 1.0
    tg = torch.ones(1, 1, 2, 2, 2)
    assert abs(cl_losses.hard_dice(lg, tg) - 0.5) < 1e-6      # class 1 perfect, class 2 absent -> (1 + 0)/2


def test_vit_unet_oracle_against_reference_fixture():
    """oracle/vit_unet.py (the checker that travels to the GPU box) reproduces the values the reference's own
    Generic_ViT_UNet gave in the build container (tests/golden/vit_unet_tiny.npz)."""
    from oracle import gen_golden, vit_unet
    gold = np.load(os.path.join(GOLD, "vit_unet_tiny.npz"))
    net = vit_unet.fill_parameters(vit_unet.Generic_ViT_UNet(1, 8, 3, 2, [16, 32, 32], [[2, 2, 2], [2, 2, 2]]), gen_golden.VIT_SEED)
    vals = gen_golden.vit_values(net)
    assert set(gold.files) == set(vals)
    for k in gold.files:
        np.testing.assert_allclose(gold[k], vals[k], rtol=2e-5, atol=2e-6, err_msg=k)


def test_augment_and_sliding_window_oracles_reproduce_their_fixtures():
    """tests/golden/augment_tiny.npz / sliding_window_tiny.npz (oracle/gen_golden_f.py) == what the oracles produce today: pins the
    scipy.ndimage-based patch-pipeline restatement and the tiled-prediction restatement; the GPU tests hold the CUDA path to the
    same files"""
    import util
    from oracle import augment as oaug, gen_golden_f
    cases, plan, patch, strides, gen_patch, params, data, targets, margin = util.load_augment_fixture()
    assert len(plan["cases"]) == 2 and all(s["angles"] is not None for s in plan["spatial"]) and gen_patch[0] > patch[0]
    d, t, m = oaug.apply_plan([c["data"] for c in cases], plan, patch, gen_patch, strides)
    np.testing.assert_allclose(d, data, rtol=1e-5, atol=1e-6)
    for a, b in zip(t, targets):
        assert np.array_equal(a, b)
    bad, agree = util.augment_mismatch(d, t, data, targets, margin)
    assert bad == 0.0 and agree == 1.0
    z = np.load(os.path.join(util.ROOT, "tests", "golden", "sliding_window_tiny.npz"))
    v = gen_golden_f.sliding_values()
    np.testing.assert_allclose(v["prob"], z["prob"], rtol=1e-4, atol=1e-6)
    assert (v["seg"] == z["seg"]).mean() > 0.9999 and np.array_equal(v["x"], z["x"])
    assert abs(float(z["prob"].sum(0).mean()) - 1) < 1e-5
