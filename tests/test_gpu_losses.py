"""GPU parity of the loss / regulariser / optimiser kernels (through the C ABI) against the oracle restatements in
oracle/cl_losses.py (which tests/test_oracle_vs_reference.py pins to the reference's own files).
Tolerance: 1e-3 relative on values and gradients (north_star); Fisher / RW updates and EWC gradients bit-stable."""
import pytest
import torch

from util import rel_err

pytestmark = pytest.mark.gpu
TOL = 1e-3


def _logits(B, C, shape, seed, scale=2.0):
    g = torch.Generator().manual_seed(seed)
    return scale * torch.randn((B, C) + shape, generator=g)


def _target(B, C, shape, seed):
    g = torch.Generator().manual_seed(seed)
    return torch.randint(0, C, (B, 1) + shape, generator=g).float()


@pytest.mark.parametrize("batch_dice", [False, True])
@pytest.mark.parametrize("C", [2, 3])
def test_dc_ce_multilevel(batch_dice, C):
    from b200unet.deep_supervision import DC_and_CE_loss, MultipleOutputLoss2
    from oracle import cl_losses
    shapes = [(8, 16, 16), (4, 8, 8), (2, 4, 4)]
    weights = cl_losses.ds_loss_weights(3)
    xs = [_logits(2, C, s, 10 + i).requires_grad_() for i, s in enumerate(shapes)]
    ys = [_target(2, C, s, 20 + i) for i, s in enumerate(shapes)]
    ref = cl_losses.multiple_output_loss2(xs, ys, weights, loss=lambda a, b: cl_losses.dc_and_ce(a, b, batch_dice=batch_dice))
    ref.backward()
    cx = [x.detach().cuda().requires_grad_() for x in xs]
    loss = MultipleOutputLoss2(DC_and_CE_loss({'batch_dice': batch_dice, 'smooth': 1e-5, 'do_bg': False}, {}), weights)
    out = loss(cx, [y.cuda() for y in ys])
    assert abs(float(out) - float(ref)) < TOL * abs(float(ref))
    out.backward()
    for a, b, w in zip(cx, xs, weights):
        if w == 0:
            assert a.grad is None and b.grad is None
        else:
            assert rel_err(a.grad, b.grad) < TOL


@pytest.mark.parametrize("B,batch_dice", [(25, False), (32, True), (17, False)])
def test_dc_ce_large_batches(B, batch_dice):
    """PLOP trains with batch 25 from the second task on (plop:85): the loss kernels take up to 32 samples per launch"""
    from b200unet.deep_supervision import DC_and_CE_loss, MultipleOutputLoss2
    from oracle import cl_losses
    x, y = _logits(B, 3, (6, 10, 12), 3).requires_grad_(), _target(B, 3, (6, 10, 12), 4)
    ref = cl_losses.dc_and_ce(x, y, batch_dice=batch_dice)
    ref.backward()
    cx = x.detach().cuda().requires_grad_()
    out = MultipleOutputLoss2(DC_and_CE_loss({'batch_dice': batch_dice, 'smooth': 1e-5, 'do_bg': False}, {}), None)([cx], [y.cuda()])
    assert abs(float(out) - float(ref)) < TOL * abs(float(ref))
    out.backward()
    assert rel_err(cx.grad, x.grad) < TOL
    with pytest.raises(Exception):
        bx = _logits(33, 3, (2, 4, 4), 5).cuda()
        MultipleOutputLoss2(DC_and_CE_loss({'batch_dice': False, 'smooth': 1e-5, 'do_bg': False}, {}), None)([bx], [_target(33, 3, (2, 4, 4), 6).cuda()])


def _param_set(seed=0):
    g = torch.Generator().manual_seed(seed)
    shapes = [(8, 1, 3, 3, 3), (8,), (16, 8, 3, 3, 3), (3, 8, 1, 1, 1), (5,), (1037,)]
    return [("p%d" % i, torch.randn(s, generator=g)) for i, s in enumerate(shapes)]


def test_ewc_and_rw_penalty():
    from b200unet import synth
    from b200unet.deep_supervision import DC_and_CE_loss, MultipleOutputLossEWC, MultipleOutputLossRW
    from oracle import cl_losses
    named = _param_set()
    fA, pA, sA = synth.make_ewc_state(named, seed=7, with_scores=True)
    fB, pB, sB = synth.make_ewc_state(named, seed=8, with_scores=True)
    fisher, params, scores = {"A": fA, "B": fB}, {"A": pA, "B": pB}, {"A": sA, "B": sB}
    x = [_logits(2, 3, (4, 8, 8), 1)]
    y = [_target(2, 3, (4, 8, 8), 2)]
    base = DC_and_CE_loss({'batch_dice': False, 'smooth': 1e-5, 'do_bg': False}, {})
    cu = lambda d: {t: {k: v.cuda() for k, v in d[t].items()} for t in d}
    for strict in (True, False):
        ops = [(n, p.clone().requires_grad_()) for n, p in named]
        cps = [(n, p.clone().cuda().requires_grad_()) for n, p in named]
        ref_pen = cl_losses.ewc_penalty(ops, fisher, params, 0.4, strict_reference=strict)
        ref_pen.backward()
        loss = MultipleOutputLossEWC(base, [1.0], 0.4, cu(fisher), cu(params), None)
        # a generator reproduces Q1 (only the first task is penalised); a list gives the documented math
        loss.update_network_params((x_ for x_ in cps) if strict else cps)
        cx = [x[0].cuda()]
        total = loss(cx, [y[0].cuda()])
        base_only = loss(cx, [y[0].cuda()], reg=False)
        pen = float(total) - float(base_only)
        assert abs(pen - float(ref_pen)) < TOL * abs(float(ref_pen))
        total.backward()
        for (n, a), (_, b) in zip(cps, ops):
            assert rel_err(a.grad, b.grad) < TOL, n
    # RW: tasks = keys[:-1]
    ops = [(n, p.clone().requires_grad_()) for n, p in named]
    cps = [(n, p.clone().cuda().requires_grad_()) for n, p in named]
    ref_pen = cl_losses.rw_penalty(ops, fisher, params, scores, 0.4, strict_reference=True)
    ref_pen.backward()
    loss = MultipleOutputLossRW(base, [1.0], 0.4, {}, {}, {}, (x_ for x_ in cps))
    loss.update_rw_params(cu(fisher), cu(params), cu(scores))
    total = loss([x[0].cuda()], [y[0].cuda()])
    base_only = MultipleOutputLossEWC(base, [1.0])([x[0].cuda()], [y[0].cuda()], reg=False)
    assert abs((float(total) - float(base_only)) - float(ref_pen)) < TOL * abs(float(ref_pen))
    total.backward()
    for (n, a), (_, b) in zip(cps, ops):
        assert rel_err(a.grad, b.grad) < TOL, n
    # second evaluation with the exhausted generator: the regulariser vanishes (Q2)
    again = loss([x[0].cuda()], [y[0].cuda()])
    assert abs(float(again) - float(base_only)) < 1e-6


def test_lwf_and_mib():
    from b200unet.deep_supervision import MultipleOutputLossMiB, lwf_distillation
    from oracle import cl_losses
    shapes = [(4, 8, 8), (2, 4, 4)]
    xs = [_logits(2, 3, s, 30 + i).requires_grad_() for i, s in enumerate(shapes)]
    ts = [_logits(2, 3, s, 40 + i) for i, s in enumerate(shapes)]
    ys = [_target(2, 3, s, 50 + i) for i, s in enumerate(shapes)]
    ref = cl_losses.lwf_distillation(xs[0], ts[0], 2.0)
    got = lwf_distillation(xs[0].detach().cuda(), ts[0].cuda(), 2.0)
    assert abs(float(got) - float(ref)) < TOL * abs(float(ref))
    weights = [2.0 / 3, 1.0 / 3]
    ref = cl_losses.mib_loss(xs, ts, ys, weights, alpha=0.9, lkd=10)
    ref.backward()
    cx = [x.detach().cuda().requires_grad_() for x in xs]
    loss = MultipleOutputLossMiB(alpha=0.9, lkd=10, weight_factors=weights)
    got = loss(cx, [t.cuda() for t in ts], [y.cuda() for y in ys])
    assert abs(float(got) - float(ref)) < TOL * abs(float(ref))
    got.backward()
    for a, b in zip(cx, xs):
        assert rel_err(a.grad, b.grad) < TOL


def test_pod_and_plop():
    from b200unet.deep_supervision import MultipleOutputLossPLOP, local_POD
    from oracle import cl_losses
    g = torch.Generator().manual_seed(5)
    layers, layers_old = {}, {}
    for i, shp in enumerate([(2, 8, 4, 16, 16), (2, 16, 2, 8, 8), (2, 3, 4, 12, 12)]):
        layers["l%d" % i] = torch.randn(shp, generator=g)
        layers_old["l%d" % i] = layers["l%d" % i] + 0.1 * torch.randn(shp, generator=g)
    for k in layers:
        ref = cl_losses.local_pod(layers[k], layers_old[k], 3)
        got = local_POD(layers[k].cuda(), layers_old[k].cuda(), 3)
        assert abs(float(got) - float(ref)) < TOL * abs(float(ref)), k
    shapes = [(4, 8, 8), (4, 6, 6)]
    weights = [2.0 / 3, 1.0 / 3]
    xs = [_logits(2, 3, s, 60 + i).requires_grad_() for i, s in enumerate(shapes)]
    xo = [_logits(2, 3, s, 70 + i, scale=6.0) for i, s in enumerate(shapes)]
    ys = [_target(2, 3, s, 80 + i) for i, s in enumerate(shapes)]
    thr = {i: torch.tensor([0.4, 0.5, 0.6]) for i in range(2)}
    ref = cl_losses.plop_loss(xs, xo, ys, weights, thr, 1.0, layers, layers_old, 1e-2, 3)
    ref.backward()
    cx = [x.detach().cuda().requires_grad_() for x in xs]
    loss = MultipleOutputLossPLOP(nr_classes=2, pod_lambda=1e-2, scales=3, weight_factors=weights)
    cu = lambda d: {k: v.cuda() for k, v in d.items()}
    loss.update_plop_params(cu(layers_old), cu(layers), {i: t.cuda() for i, t in thr.items()}, 1.0)
    got = loss(cx, [t.cuda() for t in xo], [y.cuda() for y in ys])
    assert not torch.isnan(ref)
    assert abs(float(got) - float(ref)) < TOL * abs(float(ref))
    got.backward()
    for a, b in zip(cx, xs):
        assert rel_err(a.grad, b.grad) < TOL


def test_plop_column_without_background_is_nan_like_the_reference():
    """edge case (Q10): a (b, w) column with no background voxel gives 0/0 = NaN in the reference; same here"""
    from b200unet.deep_supervision import MultipleOutputLossPLOP
    from oracle import cl_losses
    x, xo = _logits(2, 3, (2, 4, 4), 1), _logits(2, 3, (2, 4, 4), 2)
    y = _target(2, 3, (2, 4, 4), 3)
    y[0, 0, :, :, 1] = 2.0
    thr = torch.tensor([0.4, 0.5, 0.6])
    ref = cl_losses.plop_pseudo_label_loss(x, xo, y.squeeze(1), thr, 1.0)
    assert torch.isnan(ref)
    loss = MultipleOutputLossPLOP(nr_classes=2, pod_lambda=1e-2, scales=3, weight_factors=[1.0])
    lay = {"a": torch.randn(2, 4, 2, 8, 8).cuda()}
    loss.update_plop_params(lay, lay, {0: thr.cuda()}, 1.0)
    got = loss([x.cuda()], [xo.cuda()], [y.cuda()])
    assert torch.isnan(got)


def test_sgd_clip_and_fisher_rw():
    from b200unet.optim import B2SGD, fisher_square, rw_update
    from oracle import cl_losses
    named = _param_set(1)
    g = torch.Generator().manual_seed(9)
    grads = [3.0 * torch.randn(p.shape, generator=g) for _, p in named]
    ops = [p.clone().requires_grad_() for _, p in named]
    cps = [p.clone().cuda().requires_grad_() for _, p in named]
    oopt = torch.optim.SGD(ops, 1e-2, weight_decay=3e-5, momentum=0.99, nesterov=True)
    copt = B2SGD(cps, 1e-2, weight_decay=3e-5, momentum=0.99, nesterov=True)
    for it in range(3):
        for p, c, gr in zip(ops, cps, grads):
            p.grad = gr.clone() * (it + 1)
            c.grad = gr.clone().cuda() * (it + 1)
        torch.nn.utils.clip_grad_norm_(ops, 12)
        oopt.step()
        copt.clip_and_step(12)
        for p, c in zip(ops, cps):
            assert rel_err(c, p) < 1e-5
            assert rel_err(c.grad, p.grad) < 1e-5      # clip_grad_norm_ scales .grad in place
    # Fisher = grad^2 (ewc:303): exact
    F = fisher_square([c.grad for c in cps])
    for f, c in zip(F, cps):
        assert torch.equal(f, c.grad * c.grad)
    # RW update (rw:240-262)
    prev = [c.detach().clone() + 0.01 for c in cps]
    fish = [torch.rand_like(c) * 0.01 for c in cps]
    score = [torch.zeros_like(c) for c in cps]
    ref = [cl_losses.rw_update(c.detach().cpu(), c.grad.cpu(), pv.cpu(), f.cpu(), s.cpu(), 0.9) for c, pv, f, s in zip(cps, prev, fish, score)]
    rw_update(cps, prev, fish, score, alpha=0.9, have_prev=True)
    for (rf, rs), f, s, pv, c in zip(ref, fish, score, prev, cps):
        assert rel_err(f, rf) < 1e-5
        assert rel_err(s, rs) < 1e-4
        assert torch.equal(pv, c.detach())
