"""ORACLE (test infrastructure). Generates the committed fixtures under tests/golden/ by running the REFERENCE's own,
unmodified Python files (imported from /root/reference with the nnunet shim of oracle/shim on sys.path) on seeded
inputs.  Run here (the reference cannot travel to the GPU box):

    python oracle/gen_golden.py

Fixtures:
  tests/golden/cl_losses.npz  -- EWC / RW / LwF / MiB / POD / PLOP-pseudo-label values (+ selected gradients) from
                                 reference nnunet_ext/training/loss_functions/{deep_supervision,embeddings,
                                 knowledge_distillation,crossentropy}.py
  tests/golden/vit_unet_tiny.npz -- logits / gradient norms of the REFERENCE's own Generic_ViT_UNet (V1, ViT-base;
                                 generic_ViT_UNet.py + vision_transformer.py, unmodified, on the nnunet/timm shims)
                                 for the "tiny" geometry with oracle.vit_unet.fill_parameters(seed 3) weights
  tests/golden/unet_tiny.npz  -- logits / loss / gradient norms of the oracle network on the "tiny" geometry
                                 (nnunet boundary: parity unpinned by the reference, see oracle/__init__.py)
"""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = "/root/reference"
for p in (os.path.join(ROOT, "oracle", "shim"), REF, ROOT, os.path.join(ROOT, "lifelong-nnunet_b200")):
    if p not in sys.path:
        sys.path.insert(0, p)


def seeded_case():
    """The seeded inputs shared by gen_golden.py and the tests (tests/test_oracle_vs_reference.py, tests/test_golden.py)."""
    g = torch.Generator().manual_seed(2024)
    shapes = [(6, 1, 3, 3, 3), (6,), (12, 6, 3, 3, 3), (3, 6, 1, 1, 1), (7,)]
    named = [("layer%d.w" % i, torch.randn(s, generator=g)) for i, s in enumerate(shapes)]
    fisher, params, scores = {}, {}, {}
    for t in ("A", "B", "C"):
        fisher[t] = {n: torch.randn(p.shape, generator=g).pow(2) for n, p in named}
        params[t] = {n: p + 0.1 * torch.randn(p.shape, generator=g) for n, p in named}
        scores[t] = {n: torch.rand(p.shape, generator=g) for n, p in named}
    lv = [(4, 8, 8), (2, 4, 4), (4, 6, 6)]
    xs = [2.0 * torch.randn((2, 3) + s, generator=g) for s in lv]
    xo = [5.0 * torch.randn((2, 3) + s, generator=g) for s in lv]
    ys = [torch.randint(0, 3, (2, 1) + s, generator=g).float() for s in lv]
    layers, layers_old = {}, {}
    for i, shp in enumerate([(2, 4, 3, 8, 8), (2, 6, 2, 12, 12), (2, 2, 8, 8)]):
        layers["m%d" % i] = torch.randn(shp, generator=g)
        layers_old["m%d" % i] = layers["m%d" % i] + 0.2 * torch.randn(shp, generator=g)
    thr = {i: torch.tensor([0.45, 0.5, 0.55]) for i in range(3)}
    return dict(named=named, fisher=fisher, params=params, scores=scores, xs=xs, xo=xo, ys=ys, layers=layers,
                layers_old=layers_old, thr=thr, weights=[4.0 / 7, 2.0 / 7, 1.0 / 7])


def reference_values(case):
    """Evaluate the reference's own classes on the seeded case."""
    from nnunet.training.loss_functions.dice_loss import DC_and_CE_loss
    from nnunet_ext.training.loss_functions import deep_supervision as ref
    from nnunet_ext.training.loss_functions.embeddings import local_POD
    out = {}
    base = lambda: DC_and_CE_loss({'batch_dice': False, 'smooth': 1e-5, 'do_bg': False}, {})
    w = case["weights"]
    xs = [x.clone().requires_grad_() for x in case["xs"]]
    l2 = ref.MultipleOutputLoss2(base(), w)
    v = l2(xs, case["ys"])
    v.backward()
    out["base"] = v.item()
    out["base_dx0"] = xs[0].grad.numpy().copy()
    # EWC, generator (Q1) and list variants, 2 stored tasks
    fisher2 = {t: case["fisher"][t] for t in ("A", "B")}
    params2 = {t: case["params"][t] for t in ("A", "B")}
    for tag, as_gen in (("gen", True), ("list", False)):
        ps = [(n, p.clone().requires_grad_()) for n, p in case["named"]]
        loss = ref.MultipleOutputLossEWC(base(), w, 0.4, fisher2, params2, (q for q in ps) if as_gen else ps)
        v = loss([x.clone() for x in case["xs"]], case["ys"])
        v.backward()
        out["ewc_" + tag] = v.item()
        out["ewc_%s_dp2" % tag] = ps[2][1].grad.numpy().copy()
    # RW (tasks = keys[:-1]); second evaluation shows Q2
    ps = [(n, p.clone().requires_grad_()) for n, p in case["named"]]
    loss = ref.MultipleOutputLossRW(base(), w, 0.4, {}, {}, {}, (q for q in ps))
    loss.update_rw_params(case["fisher"], case["params"], case["scores"])
    v = loss([x.clone() for x in case["xs"]], case["ys"])
    v.backward()
    out["rw_first"] = v.item()
    out["rw_dp2"] = ps[2][1].grad.numpy().copy()
    out["rw_second"] = loss([x.clone() for x in case["xs"]], case["ys"]).item()
    # LwF
    loss = ref.MultipleOutputLossLWF(base(), w, [case["xs"][0]], [case["xo"][0]], 2.0)
    out["lwf"] = loss([x.clone() for x in case["xs"]], case["ys"]).item()
    # MiB
    xs = [x.clone().requires_grad_() for x in case["xs"]]
    loss = ref.MultipleOutputLossMiB(alpha=0.9, lkd=10, weight_factors=w)
    v = loss(xs, case["xo"], case["ys"])
    v.backward()
    out["mib"] = v.item()
    out["mib_dx0"] = xs[0].grad.numpy().copy()
    # POD
    for k in case["layers"]:
        out["pod_" + k] = local_POD(case["layers"][k], case["layers_old"][k], 3).item()
    five = {k: v for k, v in case["layers"].items() if v.dim() == 5}
    five_old = {k: case["layers_old"][k] for k in five}
    loss = ref.MultipleOutputLossPOD(base(), w, 1e-2, 3)
    loss.update_plop_params(five_old, five)
    out["pod_total"] = loss([x.clone() for x in case["xs"]], case["ys"]).item()
    # PLOP pseudo-label loss per level (the full forward calls .cuda(); the per-level method is device-free)
    loss = ref.MultipleOutputLossPLOP(nr_classes=2, pod_lambda=1e-2, scales=3, weight_factors=w)
    loss.update_plop_params(five_old, five, case["thr"], 1.0)
    for i in range(3):
        x = case["xs"][i].clone().requires_grad_()
        v = loss._pseudo_label_loss(x, case["xo"][i], case["ys"][i].squeeze(), idx=i)
        v.backward()
        out["plop_pseudo_%d" % i] = v.item()
        out["plop_pseudo_%d_dx" % i] = x.grad.numpy().copy()
    return out


def unet_tiny_values():
    from b200unet import synth
    from b200unet.configs import CONFIGS
    from oracle import cl_losses, step
    geom = CONFIGS["tiny"]
    net = step.build_network(geom.in_channels, geom.base_features, geom.num_classes, [list(k) for k in geom.pool])
    data, targets = synth.make_batch(geom)
    out = net(data)
    w = cl_losses.ds_loss_weights(geom.num_pool)
    l = cl_losses.multiple_output_loss2(out, targets, w)
    l.backward()
    res = {"loss": l.item(), "logits_last": out[-1].detach().numpy().copy(),
           "logits0_slice": out[0][:, :, ::4, ::8, ::8].detach().numpy().copy()}
    for n, p in net.named_parameters():
        res["gnorm/" + n] = float(p.grad.norm()) if p.grad is not None else -1.0
    return res


VIT_SEED = 3


def vit_case():
    """Seeded input / upstream gradient shared by the generator and the tests."""
    g = torch.Generator().manual_seed(77)
    x = torch.randn(2, 1, 16, 32, 32, generator=g)
    up = [torch.randn(2, 3, 16, 32, 32, generator=g), torch.randn(2, 3, 8, 16, 16, generator=g)]
    return x, up


def vit_values(net):
    """logits + gradient norms of `net` (reference class or oracle restatement) on vit_case()."""
    x, up = vit_case()
    out = net(x)
    l = sum((o * u).sum() for o, u in zip(out, up)) / 1000.0
    l.backward()
    res = {"scalar": float(l.detach()), "logits_last": out[-1].detach().numpy().copy(),
           "logits0_slice": out[0][:, :, ::4, ::8, ::8].detach().numpy().copy()}
    for n, p in net.named_parameters():
        res["gnorm/" + n] = float(p.grad.norm()) if p.grad is not None else -1.0
    res["grad_cls_token"] = net.ViT.cls_token.grad.numpy().copy()
    return res


def reference_vit_unet(vit_version='V1'):
    """The reference's own class, built as nnViTUNetTrainer.py:117-125 does, on the tiny geometry."""
    from torch import nn
    from nnunet.network_architecture.initialization import InitWeights_He
    from nnunet_ext.network_architecture.generic_ViT_UNet import Generic_ViT_UNet
    from oracle import vit_unet
    net = Generic_ViT_UNet(1, 8, 3, 2, [16, 32, 32], 2, 2, nn.Conv3d, nn.InstanceNorm3d, {'eps': 1e-5, 'affine': True},
                           nn.Dropout3d, {'p': 0, 'inplace': True}, nn.LeakyReLU, {'negative_slope': 1e-2, 'inplace': True},
                           True, False, lambda x: x, InitWeights_He(1e-2), [[2, 2, 2], [2, 2, 2]], [[3, 3, 3]] * 3,
                           False, True, True, vit_version=vit_version, vit_type='base')
    return vit_unet.fill_parameters(net, VIT_SEED)


if __name__ == "__main__":
    assert os.path.isdir(REF), "the reference is only mounted in the build container"
    torch.set_num_threads(4)
    gold = os.path.join(ROOT, "tests", "golden")
    os.makedirs(gold, exist_ok=True)
    vals = reference_values(seeded_case())
    np.savez_compressed(os.path.join(gold, "cl_losses.npz"), **vals)
    np.savez_compressed(os.path.join(gold, "unet_tiny.npz"), **unet_tiny_values())
    np.savez_compressed(os.path.join(gold, "vit_unet_tiny.npz"), **vit_values(reference_vit_unet()))
    for k, v in vals.items():
        if np.ndim(v) == 0:
            print("%-18s %.8f" % (k, v))
