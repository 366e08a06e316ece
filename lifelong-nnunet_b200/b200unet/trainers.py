"""Hot-path methods of the reference trainers, on the CUDA path.

Mirrors (names, argument meaning, side effects) of
  nnUNetTrainerMultiHead.run_iteration / initialize_optimizer_and_scheduler   reference .../multihead/nnUNetTrainerMultiHead.py:294-301,598-656
  nnUNetTrainerSequential                                                      reference .../sequential/nnUNetTrainerSequential.py:19-65
  nnUNetTrainerEWC.run_iteration / after_train                                 reference .../ewc/nnUNetTrainerEWC.py:232-310
  nnUNetTrainerRW.run_iteration / _update_f_s_values / task-end normalisation  reference .../rw/nnUNetTrainerRW.py:147-265
  nnUNetTrainerMiB.run_iteration                                               reference .../mib/nnUNetTrainerMiB.py:105-183
  nnUNetTrainerPLOP.run_iteration / register_forward_hooks                     reference .../plop/nnUNetTrainerPLOP.py:217-353
  nnUNetTrainerPOD.run_iteration                                               reference .../pod/nnUNetTrainerPOD.py:88-95
  nnUNetTrainerLWF.run_iteration                                               reference .../lwf/nnUNetTrainerLWF.py:298-370

Out of scope here (callers of the path, not the path): the epoch loop / checkpointing / data
loaders of nnunet's NetworkTrainer, the CLI, plans handling.  A trainer is therefore constructed from a
``UNetGeometry`` (what the plans file would provide) and driven by any generator yielding
``{'data': ndarray|tensor (B,C,D,H,W), 'target': [ndarray|tensor (B,1,d,h,w), ...]}`` like nnunet's augmenter.

Data-parallel training (new functionality): pass ``ddp=DataParallelGroup()``; the gradient of the
data term is all-reduced (one NCCL all-reduce of a flat arena) BEFORE the identical-on-every-rank EWC/RW penalty
gradient is added, so the regulariser is not multiplied by the world size.
"""
import copy
import math

import numpy as np
import torch

from . import deep_supervision as ds
from .generic_UNet import Generic_UNet
from .generic_ViT_UNet import Generic_ViT_UNet
from .optim import B2SGD, fisher_square, rw_update

EPSILON = 1e-8  # reference rw/nnUNetTrainerRW.py module constant


def maybe_to_torch(d):
    if isinstance(d, (list, tuple)):
        return [maybe_to_torch(i) for i in d]
    if isinstance(d, np.ndarray):
        return torch.from_numpy(d).float()
    return d


def to_cuda(data, non_blocking=True, gpu_id=0):
    if isinstance(data, (list, tuple)):
        return [i.cuda(gpu_id, non_blocking=non_blocking) for i in data]
    return data.cuda(gpu_id, non_blocking=non_blocking)


def ds_loss_weights(net_numpool):
    """reference MultiHead:1373-1383"""
    weights = np.array([1 / (2 ** i) for i in range(net_numpool)])
    mask = np.array([True] + [True if i < net_numpool - 1 else False for i in range(1, net_numpool)])
    weights[~mask] = 0
    return weights / weights.sum()


class DataParallelGroup:
    """One process per GPU; a single all-reduce (sum, then 1/N) of the flat gradient arena per step."""

    def __init__(self):
        import torch.distributed as dist
        self.dist = dist
        self.world = dist.get_world_size()
        self.rank = dist.get_rank()

    def allreduce_mean_(self, flat):
        # NCCL averages inside the collective (one pass less over the 123 MB arena); gloo (CPU tests) has no AVG
        if flat.is_cuda and self.dist.get_backend() == "nccl":
            self.dist.all_reduce(flat, op=self.dist.ReduceOp.AVG)
            return
        self.dist.all_reduce(flat, op=self.dist.ReduceOp.SUM)
        flat.mul_(1.0 / self.world)


class nnUNetTrainerMultiHead:
    def __init__(self, geometry, precision="bf16", batch_dice=False, device=None, ddp=None, initial_lr=1e-2,
                 weight_decay=3e-5, max_num_epochs=1000, seed=0, task="task_A", use_vit=False, vit_version='V1',
                 vit_type='base'):
        self.geometry = geometry
        self.use_vit, self.vit_version, self.vit_type = use_vit, vit_version, vit_type   # run_training.py --use_vit
        self.precision = precision
        self.batch_dice = batch_dice
        self.device = torch.device("cuda", torch.cuda.current_device()) if device is None else torch.device(device)
        self.ddp = ddp
        self.initial_lr, self.weight_decay, self.max_num_epochs = initial_lr, weight_decay, max_num_epochs
        self.seed, self.task = seed, task
        self.network = self.optimizer = self.loss = None
        self.epoch = 0
        self.online_eval_tp, self.online_eval_fp, self.online_eval_fn = [], [], []
        self.was_initialized = False
        # opt-in input pipelining (not in the reference, whose `to_cuda(non_blocking=True)` is stream-ordered with the
        # compute): the H2D copy of batch i+1 is issued on a copy stream before the kernels of batch i are launched.
        # It consumes the generator one batch ahead, so it is off by default (LwF draws several batches per iteration).
        self.prefetch_inputs = False
        self._prefetched, self._copy_stream = None, None

    # -- reference MultiHead:337-456 / nnViTUNetTrainer.py:101-138 ---------------------------------------------------
    def initialize_network(self):
        g = self.geometry
        torch.manual_seed(self.seed)
        kw = dict(pool_op_kernel_sizes=[list(k) for k in g.pool], conv_kernel_sizes=[[3, 3, 3]] * (g.num_pool + 1),
                  max_num_features=g.max_features)
        if self.use_vit:     # MultiHead:357 -> nnViTUNetTrainer.initialize_network (:117-125)
            self.network = Generic_ViT_UNet(g.in_channels, g.base_features, g.num_classes, g.num_pool,
                                            [int(s) for s in g.patch], vit_version=self.vit_version,
                                            vit_type=self.vit_type, **kw)
        else:
            self.network = Generic_UNet(g.in_channels, g.base_features, g.num_classes, g.num_pool, **kw)
        self.network.precision = self.precision
        self.network.to(self.device)
        self.network.inference_apply_nonlin = lambda x: torch.softmax(x, 1)

    # -- reference MultiHead:294-301 ---------------------------------------------------------------------------------
    def initialize_optimizer_and_scheduler(self):
        assert self.network is not None, "self.initialize_network must be called first"
        self.optimizer = B2SGD(self.network.parameters(), self.initial_lr, weight_decay=self.weight_decay,
                               momentum=0.99, nesterov=True)
        self.lr_scheduler = None

    def maybe_update_lr(self, epoch=None):
        """nnUNetTrainerV2 poly schedule: lr = initial_lr * (1 - ep/max)^0.9"""
        ep = self.epoch + 1 if epoch is None else epoch
        lr = self.initial_lr * (1 - ep / self.max_num_epochs) ** 0.9
        for gparam in self.optimizer.param_groups:
            gparam['lr'] = lr

    def _base_loss(self):
        """reference MultiHead:1373-1386"""
        self.ds_loss_weights = ds_loss_weights(self.geometry.num_pool)
        return ds.DC_and_CE_loss({'batch_dice': self.batch_dice, 'smooth': 1e-5, 'do_bg': False}, {})

    def initialize_loss(self):
        self.loss = ds.MultipleOutputLoss2(self._base_loss(), self.ds_loss_weights)

    def initialize(self):
        self.initialize_network()
        self.initialize_optimizer_and_scheduler()
        self.initialize_loss()
        self.was_initialized = True

    # -- gradient synchronisation hook (data parallel) ---------------------------------------------------------------
    def _sync_gradients(self):
        if self.ddp is None:
            return
        plan = getattr(self.network, "_last_plan", None)
        flat = getattr(plan, "last_flat_grad", None)
        params = [p for p in self.network.parameters() if p.grad is not None]
        if flat is not None:
            lo, hi = flat.data_ptr(), flat.data_ptr() + flat.numel() * 4
            outside = [p for p in params if not (lo <= p.grad.data_ptr() < hi)]
            if len(outside) < len(params):
                self.ddp.allreduce_mean_(flat)     # U-Net grads are views of the plan's arena: one collective
                params = outside                    # (ViT grads live in autograd-owned tensors: second bucket)
        if params:
            buf = torch.cat([p.grad.reshape(-1) for p in params])
            self.ddp.allreduce_mean_(buf)
            o = 0
            for p in params:
                p.grad.copy_(buf[o:o + p.numel()].view_as(p.grad))
                o += p.numel()

    def _forward_loss(self, data, target):
        output = self.network(data)
        return output, self.loss(output, target)

    def _backward(self, l):
        l.backward()
        self._sync_gradients()

    # -- reference MultiHead:598-656 (fp32 branch; bf16 needs no loss scaling) ----------------------------------------
    def run_iteration(self, data_generator, do_backprop=True, run_online_evaluation=False, detach=True, no_loss=False):
        data, target = self._next_batch(data_generator)

        self.optimizer.zero_grad()
        l = None
        if no_loss:
            output = self.network(data)
        else:
            output, l = self._forward_loss(data, target)
        del data
        if do_backprop and l is not None:
            self._backward(l)
            self.optimizer.clip_and_step(12)
        if run_online_evaluation:
            self.run_online_evaluation(output, target)
        del target
        if do_backprop:
            self.update_after_iteration()
        if not no_loss:
            if detach:
                l = l.detach().cpu().numpy()
            return l

    def _fetch(self, data_generator, stream=None):
        data_dict = next(data_generator)
        data = maybe_to_torch(data_dict['data'])
        target = maybe_to_torch(data_dict['target'])
        if stream is None:
            return to_cuda(data, gpu_id=self.device.index), to_cuda(target, gpu_id=self.device.index), None
        with torch.cuda.stream(stream):
            data = to_cuda(data, gpu_id=self.device.index)
            target = to_cuda(target, gpu_id=self.device.index)
            ev = torch.cuda.Event()
            ev.record(stream)
        return data, target, ev

    def _next_batch(self, data_generator):
        """reference MultiHead:606-615 (next / maybe_to_torch / to_cuda); with `prefetch_inputs` the copy of the NEXT
        batch is started on a side stream right away so that it overlaps this iteration's kernels"""
        if not self.prefetch_inputs:
            self._prefetched = None
            data, target, _ = self._fetch(data_generator)
            return data, target
        if self._copy_stream is None:
            self._copy_stream = torch.cuda.Stream(self.device)
        cur = torch.cuda.current_stream(self.device)
        if self._prefetched is not None and self._prefetched[0] is data_generator:
            _, data, target, ev = self._prefetched
        else:
            data, target, ev = self._fetch(data_generator, self._copy_stream)
        cur.wait_event(ev)
        for t in ([data] if torch.is_tensor(data) else list(data)) + ([target] if torch.is_tensor(target) else list(target)):
            t.record_stream(cur)
        self._prefetched = (data_generator,) + self._fetch(data_generator, self._copy_stream)
        return data, target

    def update_after_iteration(self):
        """reference MultiHead_Module.update_after_iteration (MultiHead_Module.py:139-157) re-splits the model and
        deep-copies the head every iteration; here the active head's parameters ARE the running model's parameters
        (shared storage), so there is nothing to copy."""
        return None

    # -- reference MultiHead:924-961 ------------------------------------------------------------------------------------
    def run_online_evaluation(self, output, target):
        import ctypes as C
        from . import _lib
        lib = _lib.load()
        out, tgt = output[0].detach().contiguous(), target[0].contiguous().float()
        B, Cc = int(out.shape[0]), int(out.shape[1])
        V = out[0, 0].numel()
        counts = torch.empty((B, Cc - 1, 3), dtype=torch.float32, device=out.device)
        scr = torch.empty(int(lib.b2_kd_scratch_bytes(B, Cc, V)), dtype=torch.uint8, device=out.device)
        _lib.check(lib.b2_online_eval(out.data_ptr(), tgt.data_ptr(), B, Cc, V, counts.data_ptr(), scr.data_ptr(),
                                      C.c_void_p(torch.cuda.current_stream(out.device).cuda_stream)))
        c = counts.cpu().numpy()
        self.online_eval_tp.append(c[..., 0])
        self.online_eval_fp.append(c[..., 1])
        self.online_eval_fn.append(c[..., 2])

    def finish_online_evaluation(self):
        """global Dice per foreground class, 2tp / (2tp + fp + fn) (reference MultiHead:1019)"""
        tp = np.sum(np.concatenate(self.online_eval_tp, 0), 0)
        fp = np.sum(np.concatenate(self.online_eval_fp, 0), 0)
        fn = np.sum(np.concatenate(self.online_eval_fn, 0), 0)
        self.online_eval_tp, self.online_eval_fp, self.online_eval_fn = [], [], []
        return [float(2 * i / (2 * i + j + k + 1e-8)) for i, j, k in zip(tp, fp, fn)]


class nnUNetTrainerSequential(nnUNetTrainerMultiHead):
    """Plain sequential fine-tuning baseline (BASELINE.json config 1): the MultiHead iteration as is."""


# ------------------------------------------------------------------------------------------------------------------------
class nnUNetTrainerEWC(nnUNetTrainerMultiHead):
    def __init__(self, *a, ewc_lambda=0.4, **kw):
        super().__init__(*a, **kw)
        self.ewc_lambda = ewc_lambda
        self.fisher, self.params = dict(), dict()

    def initialize_loss(self):
        """reference ewc:131-140"""
        self.loss = ds.MultipleOutputLossEWC(self._base_loss(), self.ds_loss_weights, self.ewc_lambda, self.fisher,
                                             self.params, self.network.named_parameters())

    def _forward_loss(self, data, target):
        # the data term goes through autograd; the penalty (value + analytic gradient) is added after backward, straight
        # into param.grad -- after the gradient all-reduce under data parallelism, because it is identical on every rank
        output = self.network(data)
        return output, self.loss(output, target, reg=False)

    def _penalty_after_backward(self):
        return self.loss.penalty_into_grads(self.ewc_lambda / 2)

    def _backward(self, l):
        l.backward()
        self._sync_gradients()
        self._last_penalty = self._penalty_after_backward()

    def run_iteration(self, data_generator, do_backprop=True, run_online_evaluation=False, detach=True, no_loss=False):
        self._last_penalty = None
        loss = super().run_iteration(data_generator, do_backprop, run_online_evaluation, False, no_loss)
        if loss is not None:
            if not do_backprop:      # validation iterations: value of the regulariser without touching gradients
                self.loss.update_network_params(self.network.named_parameters())
                loss = self.loss._penalty(loss.detach(), self.ewc_lambda / 2)
            elif self._last_penalty is not None:
                loss = loss.detach() + self._last_penalty
            if detach:
                loss = loss.detach().cpu().numpy()
        # reference ewc:247 -- a fresh generator for the next iteration
        self.loss.update_network_params(self.network.named_parameters())
        return loss

    def after_train(self, data_generator, num_batches=1):
        """reference ewc:252-310: gradients are zeroed every iteration, so only the LAST batch's squared gradient
        becomes the Fisher map; `num_batches` = 250 in the reference (249 of them are discarded work)."""
        for _ in range(num_batches):
            self.optimizer.zero_grad()
            data_dict = next(data_generator)
            data = to_cuda(maybe_to_torch(data_dict['data']), gpu_id=self.device.index)
            target = to_cuda(maybe_to_torch(data_dict['target']), gpu_id=self.device.index)
            output = self.network(data)
            loss = self.loss(output, target)
            loss.backward()
            self._sync_gradients()
        self.fisher[self.task], self.params[self.task] = dict(), dict()
        named = list(self.network.named_parameters())
        with_grad = [(n, p) for n, p in named if p.grad is not None]
        for (n, p), f in zip(with_grad, fisher_square([p.grad for _, p in with_grad])):
            self.fisher[self.task][n] = f
        for n, p in named:
            if p.grad is None:
                self.fisher[self.task][n] = torch.tensor([1], device=self.device)     # ewc:300-301
            self.params[self.task][n] = p.data.clone()
        self.loss.update_ewc_params(self.fisher, self.params)
        self.loss.update_network_params(self.network.named_parameters())


# ------------------------------------------------------------------------------------------------------------------------
class nnUNetTrainerRW(nnUNetTrainerMultiHead):
    def __init__(self, *a, rw_lambda=0.4, rw_alpha=0.9, fisher_update_after=10, **kw):
        super().__init__(*a, **kw)
        self.rw_lambda, self.alpha, self.fisher_update_after = rw_lambda, rw_alpha, fisher_update_after
        self.fisher, self.params, self.scores = dict(), dict(), dict()
        self.prev_param, self.count = None, 0
        self.finished_training_on = []

    def initialize_loss(self):
        self.loss = ds.MultipleOutputLossRW(self._base_loss(), self.ds_loss_weights, self.rw_lambda, self.fisher,
                                            self.params, self.scores, self.network.named_parameters())

    def start_task(self, task):
        """reference rw:160-168"""
        self.task = task
        self.params[task] = dict()
        self.fisher[task] = {n: torch.zeros_like(p, requires_grad=False) for n, p in self.network.named_parameters() if p.requires_grad}
        self.scores[task] = {n: torch.zeros_like(p, requires_grad=False) for n, p in self.network.named_parameters() if p.requires_grad}
        self.loss.update_rw_params(self.fisher, self.params, self.scores)
        self.loss.update_network_params(self.network.named_parameters())

    def _forward_loss(self, data, target):
        output = self.network(data)
        return output, ds.MultipleOutputLossEWC.forward(self.loss, output, target, reg=False)

    def _backward(self, l):
        l.backward()
        self._sync_gradients()
        self._last_penalty = self.loss.penalty_into_grads(self.rw_lambda, self.loss.parameter_importance)

    def run_iteration(self, data_generator, do_backprop=True, run_online_evaluation=False, detach=True, no_loss=False):
        self._last_penalty = None
        loss = super().run_iteration(data_generator, do_backprop, run_online_evaluation, False, no_loss)
        if loss is not None and self._last_penalty is not None:
            loss = loss.detach() + self._last_penalty
        self._update_f_s_values()
        if detach and loss is not None:
            loss = loss.detach().cpu().numpy()
        return loss

    def _update_f_s_values(self):
        """reference rw:231-265, one fused multi-tensor launch instead of ~10 tiny ops per tensor and a CPU round trip
        of prev_param."""
        if self.count % self.fisher_update_after == 0:
            sel = [(n, p) for n, p in self.network.named_parameters() if p.grad is not None]
            have_prev = self.prev_param is not None
            if not have_prev:
                self.prev_param = {n: torch.empty_like(p) for n, p in sel}
            rw_update([p for _, p in sel], [self.prev_param[n] for n, _ in sel], [self.fisher[self.task][n] for n, _ in sel],
                      [self.scores[self.task][n] for n, _ in sel], self.alpha, EPSILON, have_prev)
        self.count += 1

    def finish_task(self):
        """reference rw:172-200 incl. quirk Q4 (Fisher normalised with the min/max of the per-tensor score maxima)."""
        self.prev_param, self.count = None, 0
        for n, p in self.network.named_parameters():
            self.params[self.task][n] = p.data.clone()
        self.finished_training_on.append(self.task)
        values = [torch.max(v) for v in self.scores[self.task].values()]
        minim, maxim = min(values), max(values)
        for k, v in self.fisher[self.task].items():
            self.fisher[self.task][k] = (v - minim) / (maxim - minim + EPSILON)
        if len(self.finished_training_on) == 1:
            for k, v in self.scores[self.task].items():
                self.scores[self.task][k] = 2 * ((v - minim) / (maxim - minim + EPSILON))
        elif len(self.finished_training_on) > 2:
            prev = {k: v.clone() for k, v in self.scores[self.finished_training_on[-1]].items()}
            for k, v in self.scores[self.task].items():
                self.scores[self.task][k] = 0.5 * (prev[k] + (v - minim) / (maxim - minim + EPSILON))
        self.loss.update_rw_params(self.fisher, self.params, self.scores)


# ------------------------------------------------------------------------------------------------------------------------
class _TeacherMixin:
    def make_teacher(self):
        """reference mib:90-97 / plop:184-195: network_old = deepcopy(network), frozen"""
        self.network_old = copy.deepcopy(self.network)
        self.network_old._plans = {}
        for p in self.network_old.parameters():
            p.requires_grad_(False)
        self.network_old.eval()


class nnUNetTrainerMiB(nnUNetTrainerMultiHead, _TeacherMixin):
    def __init__(self, *a, mib_alpha=1., mib_lkd=10, **kw):
        super().__init__(*a, **kw)
        self.alpha, self.lkd = mib_alpha, mib_lkd
        self.network_old = None

    def initialize_loss(self):
        self._base_loss()
        self.loss_base = ds.MultipleOutputLoss2(ds.DC_and_CE_loss({'batch_dice': self.batch_dice, 'smooth': 1e-5, 'do_bg': False}, {}),
                                                self.ds_loss_weights)
        self.loss = self.loss_base
        self.MiBLoss = ds.MultipleOutputLossMiB(self.alpha, self.lkd, self.ds_loss_weights)

    def _forward_loss(self, data, target):
        output = self.network(data)
        if self.network_old is None:                       # first task: plain loss (mib:118-135)
            return output, self.loss_base(output, target)
        with torch.no_grad():                              # Q8: the reference's teacher backward is discarded work
            output_o = self.network_old(data)
        return output, self.MiBLoss(output, output_o, target)


class nnUNetTrainerPOD(nnUNetTrainerMultiHead, _TeacherMixin):
    def __init__(self, *a, pod_lambda=1e-2, scales=3, **kw):
        super().__init__(*a, **kw)
        self.pod_lambda, self.scales = pod_lambda, scales
        self.network_old = None
        self.interm_results, self.old_interm_results = dict(), dict()

    def initialize_loss(self):
        self.loss_base = ds.MultipleOutputLoss2(self._base_loss(), self.ds_loss_weights)
        self.loss = self.loss_base
        self.PODLoss = ds.MultipleOutputLossPOD(self._base_loss(), self.ds_loss_weights, self.pod_lambda, self.scales)

    def register_forward_hooks(self, old=True):
        """reference plop:330-353: a hook on every module whose type string contains 'conv.Conv'"""
        net = self.network_old if old else self.network
        store = self.old_interm_results if old else self.interm_results
        handles = []
        for name, m in net.named_modules():
            if 'conv.Conv' in str(type(m)):
                def hook(module, inp, output, name=name, store=store):
                    store[name] = output.detach()
                handles.append(m.register_forward_hook(hook))
        return handles

    def start_new_task(self):
        self.make_teacher()
        self.register_forward_hooks(old=True)
        self.register_forward_hooks(old=False)

    def _pod_layers(self):
        # local_POD needs square tiles (H == W); the reference would raise on other layers
        return ({k: v for k, v in self.old_interm_results.items() if v.dim() == 5},
                {k: v for k, v in self.interm_results.items() if v.dim() == 5})

    def _forward_loss(self, data, target):
        output = self.network(data)
        if self.network_old is None:
            return output, self.loss_base(output, target)
        with torch.no_grad():
            self.network_old(data)
        old, new = self._pod_layers()
        self.PODLoss.update_plop_params(old, new)
        return output, self.PODLoss(output, target)


class nnUNetTrainerPLOP(nnUNetTrainerPOD):
    def __init__(self, *a, **kw):
        super().__init__(*a, **kw)
        self.thresholds, self.max_entropy = dict(), None

    def initialize_loss(self):
        super().initialize_loss()
        self.PLOPLoss = ds.MultipleOutputLossPLOP(self.geometry.num_classes - 1, self.pod_lambda, self.scales, self.ds_loss_weights)

    def start_new_task(self):
        super().start_new_task()
        # quirk Q15: the reference's threshold extraction compares a list with 0, so every class falls back to
        # base_threshold = 0.001 (plop:136-173); max_entropy = log(C) (plop:122)
        C_ = self.geometry.num_classes
        self.max_entropy = torch.log(torch.tensor(C_).float()).item()
        self.thresholds = {i: torch.full((C_,), 1e-3, device=self.device) for i in range(self.geometry.num_pool)}

    def _forward_loss(self, data, target):
        output = self.network(data)
        if self.network_old is None:
            return output, self.loss_base(output, target)
        with torch.no_grad():
            output_o = self.network_old(data)
        old, new = self._pod_layers()
        self.PLOPLoss.update_plop_params(old, new, self.thresholds, self.max_entropy)
        return output, self.PLOPLoss(output, output_o, target)


class nnUNetTrainerLWF(nnUNetTrainerMultiHead):
    """LwF with the previous tasks' heads (reference lwf:298-370).  A head is the state of the split module
    (``seg_outputs`` by default, reference run_training.py:103-107); heads of finished tasks are frozen copies."""

    def __init__(self, *a, lwf_temperature=2.0, split="seg_outputs", **kw):
        super().__init__(*a, **kw)
        self.lwf_temperature, self.split = lwf_temperature, split
        self.heads = dict()          # task -> state_dict of the head sub-module
        self.target_logits = dict()  # task -> list of stored full-res logits (lwf:247-251)

    def initialize_loss(self):
        self.loss = ds.MultipleOutputLossLWF(self._base_loss(), self.ds_loss_weights, list(), list(), self.lwf_temperature)

    def _head_module(self, net=None):
        m = self.network if net is None else net
        for part in self.split.split('.'):
            m = getattr(m, part)
        return m

    def finish_task(self):
        self.heads[self.task] = {k: v.detach().clone() for k, v in self._head_module().state_dict().items()}

    def _forward_with_head(self, data, task):
        """assemble_model(task) + eval forward + first output (lwf:315-346), without touching the training state"""
        head = self._head_module()
        cur = {k: v.detach().clone() for k, v in head.state_dict().items()}
        head.load_state_dict(self.heads[task])
        with torch.no_grad():
            out = self.network(data)[0].clone()
        head.load_state_dict(cur)
        return out

    def store_target_logits(self, data_batches):
        """reference helpful_functions.calculate_target_logits (:207-266): logits of every old head on the new task's
        batches, computed once before training"""
        for task in self.heads:
            self.target_logits[task] = [self._forward_with_head(to_cuda(maybe_to_torch(d), gpu_id=self.device.index), task)
                                        for d in data_batches]

    def run_iteration(self, data_generator, do_backprop=True, run_online_evaluation=False, detach=True, no_loss=False,
                      batch_idx=0):
        if self.heads:
            # the reference intends "same batch" (Q16): peek the batch, run every old head on it, then train on it
            data_dict = next(data_generator)
            data = to_cuda(maybe_to_torch(data_dict['data']), gpu_id=self.device.index)
            preds, targets = [], []
            for task in self.heads:
                preds.append(self._forward_with_head(data, task))
                targets.append(self.target_logits[task][batch_idx % len(self.target_logits[task])])
            self.loss.update_logits(preds, targets)
            data_generator = iter([data_dict])
        return super().run_iteration(data_generator, do_backprop, run_online_evaluation, detach, no_loss)
