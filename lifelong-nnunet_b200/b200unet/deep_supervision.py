"""Drop-in loss classes with the reference's names and signatures
(reference nnunet_ext/training/loss_functions/deep_supervision.py: EWC :15-83, RW :86-135, LwF :138-214,
PLOP :217-332, POD :335-380, MiB :383-416) on top of the CUDA kernels behind include/b2unet.h.

Each ``forward`` returns a 0-dim tensor attached to autograd (trainers call ``.backward()`` on it and RW reads
``param.grad`` afterwards, reference rw:223-225).  Reference quirks that change numbers are
reproduced: Q1/Q2 (the ``network_params`` generator is consumed by the first stored task), Q5 (running division in
the POD accumulation), Q6 (POD tiling), Q10 (PLOP adaptive factor axes).
"""
import ctypes as C
import os

import torch
from torch import nn

from . import _lib


def _stream(dev):
    return C.c_void_p(torch.cuda.current_stream(dev).cuda_stream)


def _scratch(nbytes, dev):
    return torch.empty(max(int(nbytes), 256), dtype=torch.uint8, device=dev)


# --------------------------------------------------------------------------------------------------------------------
# base: MultipleOutputLoss2(DC_and_CE_loss) in two sweeps per level (value + dlogits)
# --------------------------------------------------------------------------------------------------------------------
_LEVEL_STREAMS = {}
_USE_LEVEL_STREAMS = os.environ.get("B2_LOSS_STREAMS", "1") != "0"


def _level_stream(dev, i):
    """side streams for the deep-supervision levels >= 1 (a level is three short dependent launches: the low-resolution
    levels run next to the full-resolution one instead of after it)"""
    key = (dev.index, i)
    if key not in _LEVEL_STREAMS:
        _LEVEL_STREAMS[key] = torch.cuda.Stream(dev)
    return _LEVEL_STREAMS[key]


class _DSLossFunction(torch.autograd.Function):
    @staticmethod
    def forward(ctx, cfg, targets, *logits):
        lib = _lib.load()
        dev = logits[0].device
        dls, parts, events = [], [], []
        need_grad = any(l.requires_grad for l in logits)
        cur = torch.cuda.current_stream(dev)
        # every layout / dtype conversion is enqueued on the current stream BEFORE `start` is recorded: the side streams
        # of the levels >= 1 wait on `start` only, so a conversion kernel issued later could race with their reads
        logits = [None if (float(cfg['weights'][i]) == 0 and i > 0) else x.contiguous() for i, x in enumerate(logits)]
        targets = [None if x is None else y.contiguous().float() for x, y in zip(logits, targets)]
        start = torch.cuda.Event()
        start.record(cur)
        for i, (x, y) in enumerate(zip(logits, targets)):
            w = float(cfg['weights'][i])
            if x is None:
                dls.append(None)
                continue
            B, Cc = int(x.shape[0]), int(x.shape[1])
            V = x[0, 0].numel()
            if y.numel() != B * V:
                raise ValueError("target %d has %d elements, expected %d" % (i, y.numel(), B * V))
            side = i > 0 and _USE_LEVEL_STREAMS
            stream = _level_stream(dev, i) if side else cur
            if side:
                stream.wait_event(start)
            with torch.cuda.stream(stream):
                part = torch.zeros(1, dtype=torch.float32, device=dev)
                dl = torch.empty_like(x) if need_grad else None
                scr = _scratch(lib.b2_dsloss_scratch_bytes(B, Cc, V), dev)
                _lib.check(lib.b2_dsloss_fwd_bwd(x.data_ptr(), y.data_ptr(), B, Cc, V, w, int(cfg['batch_dice']),
                                                 float(cfg['smooth']), int(cfg['do_bg']), int(cfg['ignore_index']),
                                                 int(cfg['with_dice']), None if dl is None else dl.data_ptr(),
                                                 part.data_ptr(), scr.data_ptr(), _stream(dev)))
                if side:
                    ev = torch.cuda.Event()
                    ev.record(stream)
                    events.append(ev)
                    for t in (part, dl, scr, x, y):
                        if t is not None:
                            t.record_stream(stream)
            parts.append(part)
            dls.append(dl)
        for ev in events:
            cur.wait_event(ev)
        for t in parts + [d for d in dls if d is not None]:
            t.record_stream(cur)
        ctx.dls = dls
        # fixed summation order (level 0 first), one launch
        return torch.cat(parts).sum() if len(parts) > 1 else parts[0][0]

    @staticmethod
    def backward(ctx, g):
        return (None, None) + tuple(None if d is None else d * g for d in ctx.dls)


class DC_and_CE_loss(nn.Module):
    """Mirror of nnunet's ``DC_and_CE_loss(soft_dice_kwargs, ce_kwargs)`` -- a config carrier;
    the arithmetic happens inside ``MultipleOutputLoss2`` so that all levels share one code path."""

    def __init__(self, soft_dice_kwargs, ce_kwargs, aggregate="sum", square_dice=False, weight_ce=1, weight_dice=1,
                 log_dice=False, ignore_label=None):
        super().__init__()
        if aggregate != "sum" or square_dice or log_dice or ignore_label is not None or weight_ce != 1 or weight_dice != 1:
            raise NotImplementedError("only the trainer configuration (MultiHead:1385) is supported")
        self.batch_dice = bool(soft_dice_kwargs.get('batch_dice', False))
        self.smooth = float(soft_dice_kwargs.get('smooth', 1.))
        self.do_bg = bool(soft_dice_kwargs.get('do_bg', True))
        self.ignore_index = int(ce_kwargs.get('ignore_index', -100))

    def cfg(self, weights):
        return dict(weights=weights, batch_dice=self.batch_dice, smooth=self.smooth, do_bg=self.do_bg,
                    ignore_index=self.ignore_index, with_dice=1)

    def forward(self, net_output, target):
        return _DSLossFunction.apply(self.cfg([1.0]), [target], net_output)


class RobustCrossEntropyLoss(nn.Module):
    """Mirror of reference loss_functions/crossentropy.py:18-23 (kwargs-accepting RCEL); CE only."""

    def __init__(self, **kwargs):
        super().__init__()
        self.ignore_index = int(kwargs.get('ignore_index', -100))

    def cfg(self, weights):
        return dict(weights=weights, batch_dice=0, smooth=0., do_bg=0, ignore_index=self.ignore_index, with_dice=0)

    def forward(self, net_output, target):
        return _DSLossFunction.apply(self.cfg([1.0]), [target], net_output)


class MultipleOutputLoss2(nn.Module):
    """Mirror of nnunet's ``MultipleOutputLoss2(loss, weight_factors)``."""

    def __init__(self, loss, weight_factors=None):
        super().__init__()
        self.weight_factors = weight_factors
        self.loss = loss

    def forward(self, x, y):
        assert isinstance(x, (tuple, list)), "x must be either tuple or list"
        assert isinstance(y, (tuple, list)), "y must be either tuple or list"
        weights = [1] * len(x) if self.weight_factors is None else list(self.weight_factors)
        if hasattr(self.loss, 'cfg'):
            return _DSLossFunction.apply(self.loss.cfg(weights), list(y), *x)
        l = weights[0] * self.loss(x[0], y[0])          # arbitrary user loss: plain composition
        for i in range(1, len(x)):
            if weights[i] != 0:
                l = l + weights[i] * self.loss(x[i], y[i])
        return l


# --------------------------------------------------------------------------------------------------------------------
# EWC / RW
# --------------------------------------------------------------------------------------------------------------------
class _QuadPenFunction(torch.autograd.Function):
    """value = coef * sum (F [+ S]) (theta - theta*)^2 over a list of tensors, analytic gradient, one launch."""

    @staticmethod
    def forward(ctx, coef, fishers, stars, importances, *params):
        lib = _lib.load()
        dev = params[0].device
        n = len(params)
        total = sum(p.numel() for p in params)
        need_grad = any(p.requires_grad for p in params)
        flat = torch.zeros(total, dtype=torch.float32, device=dev) if need_grad else None
        table = (_lib.PenEntry * n)()
        views, o = [], 0
        for i, p in enumerate(params):
            k = p.numel()
            for t in (fishers[i], stars[i]) + ((importances[i],) if importances is not None else ()):
                if t.device != dev or t.dtype != torch.float32 or t.numel() != k or not t.is_contiguous():
                    raise ValueError("EWC/RW state tensor %d must be contiguous fp32 on %s with %d elements" % (i, dev, k))
            table[i].theta = p.data_ptr()
            table[i].theta_star = stars[i].data_ptr()
            table[i].fisher = fishers[i].data_ptr()
            table[i].importance = importances[i].data_ptr() if importances is not None else None
            if need_grad:
                v = flat[o:o + k]
                views.append(v.view(p.shape))
                table[i].grad = v.data_ptr()
            else:
                table[i].grad = None
            table[i].numel = k
            o += k
        out = torch.zeros(1, dtype=torch.float32, device=dev)
        scr = _scratch(lib.b2_quadpen_scratch_bytes(n, total), dev)
        _lib.check(lib.b2_quadpen_fwd_bwd(table, n, float(coef), out.data_ptr(), scr.data_ptr(), _stream(dev)))
        ctx.views = views
        ctx.req = [p.requires_grad for p in params]
        return out[0]

    @staticmethod
    def backward(ctx, g):
        return (None, None, None, None) + tuple((v * g) if r else None for v, r in zip(ctx.views, ctx.req))


def _match(name, match_case, match, match_true):
    """reference deep_supervision.py:67-69"""
    return (match_case and match_true and all(m in name for m in match)) or \
           (match_case and not match_true and all(m not in name for m in match)) or (not match_case)


class MultipleOutputLossEWC(MultipleOutputLoss2):
    def __init__(self, loss, weight_factors=None, ewc_lambda=0.4, fisher=dict(), params=dict(), network_params=None,
                 match_sth=False, match=list(), match_true=True):
        super().__init__(loss, weight_factors)
        self.ewc_lambda = ewc_lambda
        self.tasks = list(fisher.keys())
        self.fisher, self.params = fisher, params
        self.network_params = network_params
        self.match_case, self.match, self.match_true = match_sth, match, match_true

    def update_ewc_params(self, fisher, params):
        self.tasks = list(fisher.keys())
        self.fisher, self.params = fisher, params

    def update_network_params(self, network_params):
        self.network_params = network_params

    def _penalty(self, loss, coef, importance=None):
        for task in self.tasks:
            # iterating the (possibly generator-valued) network_params exactly like the reference (:65-66) keeps
            # quirks Q1/Q2: a generator is exhausted by the first stored task
            sel = [(n, p) for n, p in self.network_params if _match(n, self.match_case, self.match, self.match_true)]
            if not sel:
                continue
            dev = loss.device
            fs = [self.fisher[task][n].to(dev) for n, _ in sel]
            st = [self.params[task][n].to(dev) for n, _ in sel]
            im = None if importance is None else [importance[task][n].to(dev) for n, _ in sel]
            # tensors without a Fisher map of the same size (grad None -> tensor([1]), ewc:300-301) broadcast
            for i, (n, p) in enumerate(sel):
                if fs[i].numel() != p.numel():
                    fs[i] = fs[i].float().expand_as(p).contiguous()
            loss = loss + _QuadPenFunction.apply(coef, fs, st, im, *[p for _, p in sel])
        return loss

    def forward(self, x, y, reg=True):
        loss = super().forward(x, y)
        if reg:
            loss = self._penalty(loss, self.ewc_lambda / 2)
        return loss

    @torch.no_grad()
    def penalty_into_grads(self, coef=None, importance=None):
        """Trainer fast path: evaluate the penalty of every stored task and ADD its analytic gradient straight into the
        existing ``param.grad`` buffers (one launch per task, no autograd nodes, no per-tensor torch ops).  Same task /
        generator semantics as ``forward`` (Q1/Q2).  Returns the penalty value as a 0-dim tensor."""
        lib = _lib.load()
        coef = self.ewc_lambda / 2 if coef is None else coef
        total = None
        for task in self.tasks:
            sel = [(n, p) for n, p in self.network_params if _match(n, self.match_case, self.match, self.match_true)]
            if not sel:
                continue
            dev = sel[0][1].device
            for _, p in sel:          # the penalty reaches parameters the data term does not (e.g. a zero-weight head)
                if p.grad is None:
                    p.grad = torch.zeros_like(p)
            table = (_lib.PenEntry * len(sel))()
            keep, numel = [], 0
            for i, (n, p) in enumerate(sel):
                f, st = self.fisher[task][n], self.params[task][n]
                if f.numel() != p.numel():
                    f = f.float().expand_as(p).contiguous()
                im = None if importance is None else importance[task][n]
                keep.append((f, st, im))
                table[i].theta, table[i].theta_star, table[i].fisher = p.data_ptr(), st.data_ptr(), f.data_ptr()
                table[i].importance = None if im is None else im.data_ptr()
                table[i].grad, table[i].numel = p.grad.data_ptr(), p.numel()
                numel += p.numel()
            out = torch.zeros(1, dtype=torch.float32, device=dev)
            scr = _scratch(lib.b2_quadpen_scratch_bytes(len(sel), numel), dev)
            _lib.check(lib.b2_quadpen_fwd_bwd(table, len(sel), float(coef), out.data_ptr(), scr.data_ptr(), _stream(dev)))
            total = out[0] if total is None else total + out[0]
        return total


class MultipleOutputLossRW(MultipleOutputLossEWC):
    def __init__(self, loss, weight_factors=None, ewc_lambda=0.4, fisher=dict(), params=dict(),
                 parameter_importance=dict(), network_params=None, match_sth=False, match=list(), match_true=True):
        super().__init__(loss, weight_factors, ewc_lambda, fisher, params, network_params, match_sth, match, match_true)
        self.parameter_importance = parameter_importance

    def update_rw_params(self, fisher, params, parameter_importance):
        super().update_ewc_params(fisher, params)
        self.parameter_importance = parameter_importance
        self.tasks = list(self.fisher.keys())[:-1]

    def forward(self, x, y):
        loss = super().forward(x, y, reg=False)
        return self._penalty(loss, self.ewc_lambda, self.parameter_importance)


# --------------------------------------------------------------------------------------------------------------------
# LwF
# --------------------------------------------------------------------------------------------------------------------
def lwf_distillation(pred, teacher, temperature):
    """reference deep_supervision.py:185-199 -- value only (both operands are detached, lwf:343-349)."""
    lib = _lib.load()
    dev = pred.device
    p, t = pred.detach().contiguous().float(), teacher.detach().to(dev).contiguous().float()
    B, Cc = int(p.shape[0]), int(p.shape[1])
    V = p[0, 0].numel()
    out = torch.zeros(1, dtype=torch.float32, device=dev)
    scr = _scratch(lib.b2_kd_scratch_bytes(B, Cc, V), dev)
    _lib.check(lib.b2_kd_lwf(p.data_ptr(), t.data_ptr(), B, Cc, V, float(temperature), out.data_ptr(), scr.data_ptr(),
                             _stream(dev)))
    return out[0]


class MultipleOutputLossLWF(MultipleOutputLoss2):
    def __init__(self, loss, weight_factors=None, pred_logits=list(), target_logits=list(), lwf_temperature=2.0):
        super().__init__(loss, weight_factors)
        self.pred_logits, self.target_logits = pred_logits, target_logits
        self.lwf_temperature = lwf_temperature
        self.scale = [item.size(-1) for item in self.target_logits]

    def update_logits(self, pred_logits, target_logits):
        self.pred_logits, self.target_logits = pred_logits, target_logits
        self.scale = [item.size(-1) for item in self.target_logits]

    def _distillation_loss(self, y, teacher_scores, scale):
        return lwf_distillation(y, teacher_scores, self.lwf_temperature)

    def forward(self, x, y):
        loss = super().forward(x, y)
        for idx, t_logit in enumerate(self.target_logits):
            loss = loss + self._distillation_loss(self.pred_logits[idx].to(loss.device), t_logit, self.scale[idx])
        return loss


# --------------------------------------------------------------------------------------------------------------------
# MiB
# --------------------------------------------------------------------------------------------------------------------
class _MiBKDFunction(torch.autograd.Function):
    @staticmethod
    def forward(ctx, alpha, scales, teachers, *logits):
        lib = _lib.load()
        dev = logits[0].device
        out = torch.zeros(1, dtype=torch.float32, device=dev)
        need_grad = any(l.requires_grad for l in logits)
        dls = []
        for x, t, s in zip(logits, teachers, scales):
            x, t = x.contiguous(), t.detach().to(dev).contiguous().float()
            if x.shape != t.shape:
                raise NotImplementedError("MiB KD with growing class sets is outside the reference's use (same labels per task)")
            B, Cc = int(x.shape[0]), int(x.shape[1])
            V = x[0, 0].numel()
            dl = torch.zeros_like(x) if need_grad else None
            scr = _scratch(lib.b2_kd_scratch_bytes(B, Cc, V), dev)
            _lib.check(lib.b2_kd_mib(x.data_ptr(), t.data_ptr(), B, Cc, V, float(alpha), float(s),
                                     None if dl is None else dl.data_ptr(), out.data_ptr(), scr.data_ptr(), _stream(dev)))
            dls.append(dl)
        ctx.dls = dls
        return out[0]

    @staticmethod
    def backward(ctx, g):
        return (None, None, None) + tuple(None if d is None else d * g for d in ctx.dls)


class UnbiasedKnowledgeDistillationLoss(nn.Module):
    """reference loss_functions/knowledge_distillation.py:3-32 for equal class counts, reduction='mean'."""

    def __init__(self, reduction='mean', alpha=1.):
        super().__init__()
        if reduction != 'mean':
            raise NotImplementedError("only reduction='mean' is on the hot path")
        self.alpha = alpha

    def forward(self, inputs, targets, mask=None):
        if mask is not None:
            raise NotImplementedError("mask is never passed by the trainers")
        return _MiBKDFunction.apply(self.alpha, [1.0], [targets], inputs)


class MultipleOutputLossMiB(MultipleOutputLoss2):
    def __init__(self, alpha=1., lkd=10, weight_factors=None):
        super().__init__(RobustCrossEntropyLoss(ignore_index=255), weight_factors)
        self.lkd, self.alpha = lkd, alpha
        self.lkd_loss = UnbiasedKnowledgeDistillationLoss(alpha=self.alpha)

    def forward(self, x, x_o, y):
        assert isinstance(x_o, (tuple, list)), "x_o must be either tuple or list"
        loss = super().forward(x, y)
        weights = self.weight_factors if self.weight_factors is not None else [1] * len(x)
        scales = [float(weights[i]) * self.lkd for i in range(len(x))]
        return loss + _MiBKDFunction.apply(self.alpha, scales, list(x_o), *x)


# --------------------------------------------------------------------------------------------------------------------
# POD / PLOP
# --------------------------------------------------------------------------------------------------------------------
def _act_view(t):
    """Describe a (B,C,D,H,W) tensor with channels-last-3d strides (what Generic_UNet's hooks hand out) -- or any dense
    tensor, which is converted -- as a b2_act_view."""
    if t.dim() == 4:
        t = t.unsqueeze(2)
    B, Cc, D, H, W = t.shape
    st = t.stride()
    ok = st[1] == 1 and st[4] >= Cc and st[3] == W * st[4] and st[2] == H * st[3] and st[0] == D * st[2]
    if not ok or t.dtype not in (torch.float32, torch.bfloat16):
        t = t.float().contiguous(memory_format=torch.channels_last_3d)
        st = t.stride()
    v = _lib.ActView()
    v.ptr, v.n, v.d, v.h, v.w, v.c, v.pitch = t.data_ptr(), B, D, H, W, Cc, st[4]
    v.dtype = _lib.B2_F32 if t.dtype == torch.float32 else _lib.B2_BF16
    return v, t


def local_POD(h_, h_old, scales):
    """reference loss_functions/embeddings.py:9-42 (value only)."""
    assert h_.size() == h_old.size(), "The embedding tensors of the current and old model should have the same shape.."
    lib = _lib.load()
    dev = h_.device
    h_old = h_old.to(dev)
    if h_old.dtype != h_.dtype:
        h_old = h_old.to(h_.dtype)
    va, ka = _act_view(h_.detach())
    vb, kb = _act_view(h_old.detach())
    out = torch.zeros(1, dtype=torch.float32, device=dev)
    scr = _scratch(lib.b2_pod_scratch_bytes(C.byref(va), int(scales)), dev)
    _lib.check(lib.b2_pod_local(C.byref(va), C.byref(vb), int(scales), out.data_ptr(), scr.data_ptr(), _stream(dev)))
    del ka, kb
    return out[0]


def _pod_running(interm, old_interm, pod_lambda, scales):
    """reference deep_supervision.py:270-278 / :368-376 (Q5: running division inside the loop)."""
    num_layers = len(old_interm.keys())
    dist = 0
    for name, h_old in old_interm.items():
        dist = dist + pod_lambda * local_POD(interm[name], h_old, scales)
        dist = dist / num_layers
    return dist


class MultipleOutputLossPOD(MultipleOutputLoss2):
    def __init__(self, loss, weight_factors=None, pod_lambda=1e-2, scales=3):
        super().__init__(loss, weight_factors)
        self.pod_lambda, self.scales = pod_lambda, scales

    def update_plop_params(self, old_interm_results, interm_results):
        self.old_interm_results, self.interm_results = old_interm_results, interm_results
        self.num_layers = len(self.old_interm_results.keys())

    def forward(self, x, y):
        loss = super().forward(x, y)
        return loss + _pod_running(self.interm_results, self.old_interm_results, self.pod_lambda, self.scales)


class _PlopPseudoFunction(torch.autograd.Function):
    @staticmethod
    def forward(ctx, weights, thresholds, max_entropy, teachers, targets, *logits):
        lib = _lib.load()
        dev = logits[0].device
        out = torch.zeros(1, dtype=torch.float32, device=dev)
        need_grad = any(l.requires_grad for l in logits)
        dls = []
        for i, (x, xo, y) in enumerate(zip(logits, teachers, targets)):
            w = float(weights[i])
            if w == 0 and i > 0:
                dls.append(None)
                continue
            x, xo = x.contiguous(), xo.detach().to(dev).contiguous().float()
            y = y.contiguous().float()
            if x.dim() != 5:
                raise NotImplementedError("PLOP pseudo-label kernel covers the 3D trainers")
            if x.shape[0] < 2:
                raise IndexError("too many indices for tensor of dimension 3 (reference PLOP needs B >= 2 on 3D data, Q10)")
            B, Cc, D, H, W = (int(s) for s in x.shape)
            thr = thresholds[i].to(dev).contiguous().float()
            dl = torch.empty_like(x) if need_grad else None
            scr = _scratch(lib.b2_kd_scratch_bytes(B, Cc, D * H * W), dev)
            _lib.check(lib.b2_plop_pseudo(x.data_ptr(), xo.data_ptr(), y.data_ptr(), B, Cc, D, H, W, thr.data_ptr(),
                                          float(max_entropy), w, None if dl is None else dl.data_ptr(), out.data_ptr(),
                                          scr.data_ptr(), _stream(dev)))
            dls.append(dl)
        ctx.dls = dls
        return out[0]

    @staticmethod
    def backward(ctx, g):
        return (None, None, None, None, None) + tuple(None if d is None else d * g for d in ctx.dls)


class MultipleOutputLossPLOP(nn.Module):
    def __init__(self, nr_classes=1, pod_lambda=1e-2, scales=3, weight_factors=None):
        super().__init__()
        self.scales, self.nr_classes, self.pod_lambda, self.weight_factors = scales, nr_classes, pod_lambda, weight_factors
        self.ce = RobustCrossEntropyLoss(ignore_index=255)

    def update_plop_params(self, old_interm_results, interm_results, thresholds, max_entropy):
        self.thresholds, self.max_entropy = thresholds, max_entropy
        self.interm_results, self.old_interm_results = interm_results, old_interm_results
        self.num_layers = len(self.old_interm_results.keys())

    def forward(self, x, x_o, y):
        assert isinstance(x, (tuple, list)), "x must be either tuple or list"
        assert isinstance(x_o, (tuple, list)), "x_o must be either tuple or list"
        assert isinstance(y, (tuple, list)), "y must be either tuple or list"
        weights = [1] * len(x) if self.weight_factors is None else list(self.weight_factors)
        pseudo = _PlopPseudoFunction.apply(weights, self.thresholds, self.max_entropy, list(x_o), list(y), *x)
        dist = _pod_running(self.interm_results, self.old_interm_results, self.pod_lambda, self.scales)
        del self.thresholds, self.max_entropy, self.interm_results, self.old_interm_results
        return pseudo + dist
