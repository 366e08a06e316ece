"""GPU parity, SURVEY.md section 8 row a2: ``Generic_ViT_UNet`` V1 -- encoder / decoder through the C ABI's partial
passes (b2_unet_forward_parts / b2_unet_backward_parts), ViT through ATen on the device -- against oracle/vit_unet.py
(pinned to the reference's own class, tests/golden/vit_unet_tiny.npz) on identical weights and inputs.

Tolerances: fp32 mode 1e-3 relative (north_star); a LeakyReLU branch flip vs the oracle (pre-activation within rounding
of 0) relaxes the gradient bound to 5e-2 as in tests/test_gpu_unet.py.  bf16 mode: cosine similarity."""
import numpy as np
import os
import pytest
import torch

import util
from util import rel_err

pytestmark = pytest.mark.gpu
TOL, TOL_FLIPPED = 1e-3, 5e-2


def _pair(precision="fp32"):
    from b200unet.generic_ViT_UNet import Generic_ViT_UNet
    from oracle import gen_golden, vit_unet
    onet = vit_unet.fill_parameters(vit_unet.Generic_ViT_UNet(1, 8, 3, 2, [16, 32, 32], [[2, 2, 2], [2, 2, 2]]), gen_golden.VIT_SEED)
    cnet = Generic_ViT_UNet(1, 8, 3, 2, [16, 32, 32], pool_op_kernel_sizes=[[2, 2, 2], [2, 2, 2]], conv_kernel_sizes=[[3, 3, 3]] * 3)
    cnet.load_state_dict(onet.state_dict())
    cnet.precision = precision
    return onet, cnet.cuda(), gen_golden.vit_case()


def test_forward_matches_oracle_and_reference_fixture():
    onet, cnet, (x, up) = _pair()
    with torch.no_grad():
        oo = onet(x)
        co = cnet(x.cuda())
    assert len(co) == 2
    for a, b in zip(co, oo):
        assert a.shape == b.shape and rel_err(a, b) < TOL
    gold = np.load(os.path.join(util.ROOT, "tests", "golden", "vit_unet_tiny.npz"))
    assert rel_err(co[1], torch.from_numpy(gold["logits_last"])) < TOL
    assert rel_err(co[0][:, :, ::4, ::8, ::8], torch.from_numpy(gold["logits0_slice"])) < TOL


def test_backward_matches_oracle():
    onet, cnet, (x, up) = _pair()
    oo = onet(x)
    (sum((o * u).sum() for o, u in zip(oo, up)) / 1000.0).backward()
    co = cnet(x.cuda())
    (sum((o * u.cuda()).sum() for o, u in zip(co, up)) / 1000.0).backward()
    # LeakyReLU branch flips in the encoder / decoder activations
    ys = []
    hooks = [m.register_forward_hook(lambda mod, i, o: ys.append(o.detach().clone()))
             for n, m in onet.named_modules() if n.endswith("lrelu")]
    with torch.no_grad():
        onet(x)
    for h in hooks:
        h.remove()
    from b200unet.generic_ViT_UNet import _view
    flips = 0
    for i, yo in enumerate(ys):
        if i in (4, 5):
            continue         # bottleneck blocks: not executed by the CUDA path (result discarded in V1)
        flips += int(((_view(cnet._last_plan, i, 1).cpu() > 0) != (yo > 0)).sum())
    tol = TOL if flips == 0 else TOL_FLIPPED
    od = dict(onet.named_parameters())
    report, bad, n_none = [], [], 0
    for n, p in cnet.named_parameters():
        ref = od[n].grad
        if ref is None:
            assert p.grad is None, n
            n_none += 1
            continue
        assert p.grad is not None, n
        scale = max(float(ref.abs().max()), 1e-6)
        if "conv.bias" in n and "seg" not in n:
            scale = max(float(od[n.replace("bias", "weight")].grad.abs().max()), 1e-6)
        err = float((p.grad.cpu() - ref).abs().max()) / scale
        report.append("%-60s %.3e" % (n, err))
        if not err < tol:
            bad.append(n)
    assert n_none == 8          # the two bottleneck blocks (conv w/b, instnorm w/b) get grad None
    assert not bad, "flips=%d\n" % flips + "\n".join(report)


def test_store_vit_input_and_attention_weights():
    onet, cnet, (x, up) = _pair()
    cnet.ViT.store_attn_weights = True
    with torch.no_grad():
        cnet(x.cuda(), store_vit_input=True)
        skips = []
        h = onet.conv_blocks_context[0].register_forward_hook(lambda m, i, o: skips.append(o))
        onet(x)
        h.remove()
    assert rel_err(cnet.ViT_in.float(), skips[0]) < TOL
    assert len(cnet.ViT.attn_weights) == 12 and cnet.ViT.attn_weights[0].shape == (2, 12, 5, 5)
    for a, b in zip(cnet.ViT.attn_weights, onet.ViT.attn_weights):
        assert rel_err(a, b) < TOL


def test_bf16_mode_close_to_oracle():
    onet, cnet, (x, up) = _pair("bf16")
    oo = onet(x)
    (sum((o * u).sum() for o, u in zip(oo, up)) / 1000.0).backward()
    co = cnet(x.cuda())
    (sum((o * u.cuda()).sum() for o, u in zip(co, up)) / 1000.0).backward()
    cos = lambda a, b: float(torch.nn.functional.cosine_similarity(a.flatten().double().cpu(), b.flatten().double(), dim=0))
    for a, b in zip(co, oo):
        assert cos(a, b) > 0.99
    od = dict(onet.named_parameters())
    for n in ("ViT.heads.0.weight", "ViT.patch_embeds.0.proj.weight", "ViT.blocks.layer.0.attn.qkv.weight",
              "conv_blocks_context.0.blocks.0.conv.weight", "tu.0.weight", "seg_outputs.1.weight"):
        assert cos(dict(cnet.named_parameters())[n].grad, od[n].grad) > 0.9, n


def test_trainer_step_with_vit():
    """one EWC-free SGD step of the MultiHead trainer with use_vit: parameters after the step match the oracle's"""
    from b200unet import synth
    from b200unet.configs import CONFIGS
    from b200unet.trainers import nnUNetTrainerMultiHead
    from oracle import cl_losses, gen_golden, step, vit_unet
    geom = CONFIGS["tiny"]
    tr = nnUNetTrainerMultiHead(geom, precision="fp32", use_vit=True)
    tr.initialize()
    onet = vit_unet.fill_parameters(vit_unet.Generic_ViT_UNet(1, 8, 3, 2, [16, 32, 32], [[2, 2, 2], [2, 2, 2]]), gen_golden.VIT_SEED)
    tr.network.load_state_dict(onet.state_dict())
    data, targets = synth.make_batch(geom, seed=5)
    w = cl_losses.ds_loss_weights(geom.num_pool)
    opt = step.make_optimizer(onet)
    ol, _ = step.run_iteration(onet, opt, data, targets, lambda o, t: cl_losses.multiple_output_loss2(o, t, w))
    gen = iter([{"data": data, "target": targets}])
    cl = tr.run_iteration(gen)
    assert abs(float(cl) - ol) < TOL * abs(ol)
    od = dict(onet.named_parameters())
    for n, p in tr.network.named_parameters():
        assert rel_err(p, od[n]) < 2e-2 if "conv.bias" in n else rel_err(p, od[n]) < 5e-3, n


@pytest.mark.parametrize("precision", ["fp32", "bf16"])
def test_trainer_with_task_specific_layernorms(precision):
    """ViT task-specific LayerNorms (vision_transformer.py:380-416, MultiHead:554-561) on the CUDA path: a second task
    registers a fresh LN set, trains only that set (the first task's LNs receive no gradient and stay bit-identical), and
    switching back selects the first set again.  The loss of the first iteration on task B equals the loss the same weights
    give with task A's (identical, freshly initialised) LNs: fresh LNs are (1, 0) like A's untouched ones."""
    from b200unet import synth
    from b200unet.configs import CONFIGS
    from b200unet.trainers import nnUNetTrainerMultiHead
    geom = CONFIGS["tiny32"] if precision == "bf16" else CONFIGS["tiny"]
    tr = nnUNetTrainerMultiHead(geom, precision=precision, use_vit=True, ViT_task_specific_ln=True, task="A", cuda_graph=False)
    tr.initialize()
    vit = tr.network.ViT
    assert vit.task_name_use == "A" and list(vit.norm.keys()) == ["A"]
    batch = lambda s: iter([dict(zip(("data", "target"), synth.make_batch(geom, seed=s)))])
    la0 = float(tr.run_iteration(batch(1), do_backprop=False))
    tr.start_task("B")
    assert list(vit.norm.keys()) == ["A", "B"] and vit.task_name_use == "B"
    lb0 = float(tr.run_iteration(batch(1), do_backprop=False))
    assert abs(la0 - lb0) <= 1e-6 * abs(la0)
    a_before = {n: p.detach().clone() for n, p in vit.named_parameters() if ".A." in n}
    b_before = {n: p.detach().clone() for n, p in vit.named_parameters() if ".B." in n}
    for s in range(3):
        tr.run_iteration(batch(2 + s))
    torch.cuda.synchronize()
    cur = dict(vit.named_parameters())
    assert all(torch.equal(cur[n], v) for n, v in a_before.items())
    assert sum(int(not torch.equal(cur[n], v)) for n, v in b_before.items()) > len(b_before) // 2
    tr.start_task("A")
    assert vit.task_name_use == "A" and vit.blocks.layer[0].use_task_name == "A"
    assert np.isfinite(float(tr.run_iteration(batch(9), do_backprop=False)))


def _pair_v(version, precision):
    from b200unet.generic_ViT_UNet import Generic_ViT_UNet
    from oracle import gen_golden, vit_unet
    onet = vit_unet.fill_parameters(vit_unet.Generic_ViT_UNet(1, 8, 3, 2, [16, 32, 32], [[2, 2, 2], [2, 2, 2]], vit_version=version),
                                    gen_golden.VIT_SEED)
    cnet = Generic_ViT_UNet(1, 8, 3, 2, [16, 32, 32], pool_op_kernel_sizes=[[2, 2, 2], [2, 2, 2]], conv_kernel_sizes=[[3, 3, 3]] * 3,
                            vit_version=version)
    cnet.load_state_dict(onet.state_dict())
    cnet.precision = precision
    return onet, cnet.cuda(), gen_golden.vit_case()


@pytest.mark.parametrize("version", ["V2", "V3"])
def test_v2_v3_forward_backward_match_oracle(version):
    """Generic_ViT_UNet V2 / V3 (generic_ViT_UNet.py:290-338; oracle pinned to the reference class in
    tests/test_oracle_vs_reference.py): the ViT input is fused from the first skip and the bottleneck (V3: and every skip)
    up-sampled through the `tu` chain -- b2_tconv3d_fwd / _bwd outside the plan.  fp32: logits 1e-3; gradients of the parameters
    the new data flow touches (tu weights used twice, live bottleneck convolutions, ViT) 5e-3 / cosine."""
    onet, cnet, (x, up) = _pair_v(version, "fp32")
    oo = onet(x)
    (sum((o * u).sum() for o, u in zip(oo, up)) / 1000.0).backward()
    co = cnet(x.cuda())
    for a, b in zip(co, oo):
        assert a.shape == b.shape and rel_err(a, b) < TOL, (version, rel_err(a, b))
    (sum((o * u.cuda()).sum() for o, u in zip(co, up)) / 1000.0).backward()
    od, cd = dict(onet.named_parameters()), dict(cnet.named_parameters())
    cos = lambda a, b: float(torch.nn.functional.cosine_similarity(a.flatten().double().cpu(), b.flatten().double(), dim=0))
    for n in ("tu.0.weight", "tu.1.weight", "conv_blocks_context.2.0.blocks.0.conv.weight", "conv_blocks_context.2.1.blocks.0.conv.weight",
              "ViT.patch_embeds.0.proj.weight", "ViT.heads.0.weight", "ViT.blocks.layer.5.mlp.fc1.weight",
              "conv_blocks_context.0.blocks.0.conv.weight", "conv_blocks_context.1.blocks.1.conv.weight", "seg_outputs.1.weight"):
        assert cd[n].grad is not None and od[n].grad is not None, n
        assert cos(cd[n].grad, od[n].grad) > 0.999 and rel_err(cd[n].grad, od[n].grad) < 5e-2, (n, rel_err(cd[n].grad, od[n].grad))
    assert rel_err(cd["tu.0.weight"].grad, od["tu.0.weight"].grad) < 5e-3
    assert rel_err(cd["conv_blocks_context.2.1.blocks.0.conv.weight"].grad, od["conv_blocks_context.2.1.blocks.0.conv.weight"].grad) < 5e-3


@pytest.mark.parametrize("version", ["V2", "V3"])
def test_v2_v3_bf16_native_vit(version):
    """production precision: native ViT kernels on the fused input, input gradient handed back to autograd for the tu chain"""
    onet, cnet, (x, up) = _pair_v(version, "bf16")
    oo = onet(x)
    (sum((o * u).sum() for o, u in zip(oo, up)) / 1000.0).backward()
    co = cnet(x.cuda())
    (sum((o * u.cuda()).sum() for o, u in zip(co, up)) / 1000.0).backward()
    cos = lambda a, b: float(torch.nn.functional.cosine_similarity(a.flatten().double().cpu(), b.flatten().double(), dim=0))
    for a, b in zip(co, oo):
        assert cos(a, b) > 0.99
    od, cd = dict(onet.named_parameters()), dict(cnet.named_parameters())
    for n in ("tu.0.weight", "tu.1.weight", "conv_blocks_context.2.1.blocks.0.conv.weight", "ViT.heads.0.weight",
              "ViT.patch_embeds.0.proj.weight", "conv_blocks_context.0.blocks.0.conv.weight"):
        assert cd[n].grad is not None and cos(cd[n].grad, od[n].grad) > 0.9, (n, cos(cd[n].grad, od[n].grad))
