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
import os

import numpy as np
import torch

from . import deep_supervision as ds
from .fused_step import FusedStep
from .generic_UNet import Generic_UNet
from .generic_ViT_UNet import Generic_ViT_UNet
from .MultiHead_Module import MultiHead_Module
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
                 vit_type='base', split="seg_outputs", transfer_heads=True, strict_reference=True, fused_step=True,
                 cuda_graph=None, ViT_task_specific_ln=False, do_LSA=False):
        self.geometry = geometry
        self.split, self.transfer_heads = split, transfer_heads       # run_training.py -s / --transfer_heads
        # strict_reference: reproduce the reference's generator-exhaustion quirks Q1 / Q2; False = the
        # documented math (every stored task is penalised on every iteration)
        self.strict_reference = strict_reference
        # fused_step: run training iterations as one enqueue program over the C ABI (fused_step.py) instead of through
        # autograd; cuda_graph: capture that program after two warm-up iterations and replay it (env B2_CUDA_GRAPH=0/1)
        self.fused_step = fused_step
        if cuda_graph is None:
            cuda_graph = os.environ.get("B2_CUDA_GRAPH", "1") != "0"
        self.cuda_graph = cuda_graph
        self._steps, self._step_inputs = {}, None
        self.mh_network = None
        self.use_vit, self.vit_version, self.vit_type = use_vit, vit_version, vit_type   # run_training.py --use_vit
        self.ViT_task_specific_ln = bool(ViT_task_specific_ln)                           # run_training.py --task_specific_ln
        self.do_LSA = bool(do_LSA)                                                       # run_training.py --do_LSA
        self.precision = precision
        self.batch_dice = batch_dice
        if device is None:
            device = torch.device("cuda", torch.cuda.current_device()) if torch.cuda.is_available() else torch.device("cpu")
        self.device = torch.device(device)
        self.extension = getattr(type(self), "EXTENSION", "multihead")      # utilities/ext_map.py
        self.fold = 0
        self.already_trained_on = {"0": {"finished_training_on": [], "start_training_on": None, "fisher_at": None,
                                         "params_at": None, "scores_at": None, "checkpoint_should_exist": False,
                                         "tasks_at_time_of_checkpoint": [], "active_task_at_time_of_checkpoint": None}}
        self._init_kwargs = dict(precision=precision, batch_dice=batch_dice, initial_lr=initial_lr, weight_decay=weight_decay,
                                 max_num_epochs=max_num_epochs, seed=seed, task=task, use_vit=use_vit, vit_version=vit_version,
                                 vit_type=vit_type, split=split, transfer_heads=transfer_heads, strict_reference=strict_reference,
                                 ViT_task_specific_ln=ViT_task_specific_ln, do_LSA=do_LSA)
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
                                            vit_type=self.vit_type, ViT_task_specific_ln=self.ViT_task_specific_ln,
                                            first_task_name=self.task, do_LSA=self.do_LSA, **kw)
            if self.ViT_task_specific_ln:    # nnViTUNetTrainer.py:128-129
                self.network.ViT.use_task(self.task)
        else:
            self.network = Generic_UNet(g.in_channels, g.base_features, g.num_classes, g.num_pool, **kw)
        self.network.precision = self.precision
        self.network.to(self.device)
        self.network.inference_apply_nonlin = lambda x: torch.softmax(x, 1)
        # reference MultiHead:366-369: the body / head splitter around the network; the running model IS self.network
        self.mh_network = MultiHead_Module(type(self.network), self.split, self.task, prev_trainer=self.network)
        self.network = self.mh_network.model

    # -- checkpoints (reference MultiHead:1164-1313; formats in checkpoint.py) ------------------------------------------------
    def init_args(self):
        return dict(geometry=self.geometry, **self._init_kwargs)

    def save_checkpoint(self, fname, save_optimizer=True):
        from . import checkpoint
        checkpoint.save_checkpoint(self, fname, save_optimizer)

    def load_checkpoint(self, fname, train=True):
        from . import checkpoint
        return checkpoint.load_checkpoint(self, fname, train)

    def finish_training_on(self, task=None):
        """bookkeeping of MultiHead.run_training :575-590"""
        t = self.task if task is None else task
        ft = self.already_trained_on[str(self.fold)]["finished_training_on"]
        if t not in ft:
            ft.append(t)

    def start_task(self, task):
        """reference MultiHead.run_training :541-564: a new task gets a head (initialised from the first split, or from
        the last trained head with transfer_heads) and the running model is assembled for it"""
        if task not in self.mh_network.heads:
            self.mh_network.add_new_task(task, use_init=not self.transfer_heads)
        new_lns = False
        if self.use_vit and self.ViT_task_specific_ln:      # MultiHead:554-561
            if task not in self.network.ViT.norm:
                self.network.ViT.register_new_task(task)
                new_lns = True
            self.network.ViT.use_task(task)
        self.network = self.mh_network.assemble_model(task)
        self.task = task
        self.already_trained_on[str(self.fold)]["start_training_on"] = task
        self._steps = {}
        if new_lns and self.optimizer is not None:          # the new LayerNorms are trainable parameters of this task
            known = {id(p) for g in self.optimizer.param_groups for p in g['params']}
            fresh = [p for p in self.network.parameters() if id(p) not in known]
            if fresh:
                self.optimizer.param_groups[0]['params'].extend(fresh)

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
    # -- fused iteration (fused_step.py) ----------------------------------------------------------------------------------
    def _matched_param_names(self):
        """parameter names the loss's name filter selects, walked once per (network, filter) -- not per iteration: the walk over
        named_parameters() costs ~0.25 ms of host time, which sits between two steps while the GPU idles.  The cache is tied to
        the identity of the self._steps dict, so everything that invalidates the step programs by replacing that dict (start_task,
        new LayerNorm sets, checkpoint loading, ...) invalidates it too."""
        key = (id(self.network), self.loss.match_case, tuple(self.loss.match or ()), self.loss.match_true)
        c = getattr(self, "_names_cache", None)
        if c is None or c[0] is not self._steps or c[1] != key:
            names = [n for n, _ in self.network.named_parameters()
                     if ds._match(n, self.loss.match_case, self.loss.match, self.loss.match_true)]
            c = self._names_cache = (self._steps, key, names)
        return c[2]

    def _use_fused(self, do_backprop, no_loss):
        return (self.fused_step and do_backprop and not no_loss and not self.use_vit and
                type(self.network) is Generic_UNet and self._fused_supported())

    def _fused_supported(self):
        return type(self.loss) is ds.MultipleOutputLoss2 and hasattr(self.loss.loss, 'cfg')

    def _fused_spec(self):
        """(hashable key, spec dict) of the iteration's loss composition; the key changes when a program must be rebuilt"""
        return ("base",), dict(base='dcce', cfg=self.loss.loss.cfg(list(self.ds_loss_weights)))

    def _fused_pre(self, step):
        pass

    def _fused_post(self, step):
        pass

    def _host_batch(self, data_generator):
        data_dict = next(data_generator)
        return maybe_to_torch(data_dict['data']), maybe_to_torch(data_dict['target'])

    def _run_iteration_fused(self, data_generator, run_online_evaluation=False, detach=True):
        """reference MultiHead:606-656 with the body (zero_grad ... optimizer.step) replaced by one FusedStep program"""
        dev = self.device
        cur = torch.cuda.current_stream(dev)
        pending = self._prefetched if (self.prefetch_inputs and self._prefetched is not None and
                                       self._prefetched[0] is data_generator) else None
        if pending is None:
            data, target = self._host_batch(data_generator)
        else:
            data, target = pending[1], pending[2]
        key, spec = self._fused_spec()
        key = (tuple(data.shape), self.precision) + key
        step = self._steps.get(key)
        if step is None:
            step = self._steps[key] = FusedStep(self, torch.empty(tuple(data.shape), dtype=torch.float32, device=dev),
                                                [torch.empty(tuple(t.shape), device='meta') for t in target], spec)
        if self._copy_stream is None:
            self._copy_stream = torch.cuda.Stream(dev)
        cs = self._copy_stream
        if pending is None:
            # input patch on the compute stream (the forward needs it first); the deep-supervision targets follow on the copy
            # stream while the forward runs and are awaited right before the loss kernels
            step.data.copy_(data, non_blocking=True)
            cs.wait_stream(cur)             # the previous iteration may still read step.targets
            with torch.cuda.stream(cs):
                for dst, src in zip(step.targets, target):
                    dst.copy_(src, non_blocking=True)
                tready = torch.cuda.Event()
                tready.record(cs)
        else:
            cur.wait_event(pending[3])      # staged on the device by the previous iteration's prefetch
            step.data.copy_(data)
            data.record_stream(cur)
            for dst, src in zip(step.targets, target):
                dst.copy_(src)
                src.record_stream(cur)
            tready = None
        self._fused_pre(step)
        total = step.run(self.cuda_graph, tready)
        self.network._last_plan = step.plan
        if self.prefetch_inputs:            # opt-in: draw batch i+1 now and stage it on the device while batch i computes
            try:
                nd, nt = self._host_batch(data_generator)
                with torch.cuda.stream(cs):
                    nd = nd.to(dev, non_blocking=True)
                    nt = [t.to(dev, non_blocking=True) for t in nt]
                    ev = torch.cuda.Event()
                    ev.record(cs)
                self._prefetched = (data_generator, nd, nt, ev)
            except StopIteration:
                self._prefetched = None
        self._fused_post(step)
        if run_online_evaluation:
            self.run_online_evaluation(step.logits, step.targets)
        self.update_after_iteration()
        if not detach:
            return total.clone()
        if getattr(self, "_loss_host", None) is None:
            self._loss_host = torch.zeros((), dtype=torch.float32).pin_memory()
        self._loss_host.copy_(total, non_blocking=True)
        cur.synchronize()
        return self._loss_host.numpy().copy()

    def run_iteration(self, data_generator, do_backprop=True, run_online_evaluation=False, detach=True, no_loss=False):
        if self._use_fused(do_backprop, no_loss):
            return self._run_iteration_fused(data_generator, run_online_evaluation, detach)
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
        try:
            self._prefetched = (data_generator,) + self._fetch(data_generator, self._copy_stream)
        except StopIteration:        # finite generator: the batch in hand is the last one -- train on it
            self._prefetched = None
        return data, target

    def update_after_iteration(self):
        """reference MultiHead_Module.update_after_iteration (MultiHead_Module.py:139-157) re-splits the model and
        deep-copies the head every iteration; here the active head's parameters ARE the running model's parameters
        (shared storage), so there is nothing to copy."""
        if self.mh_network is not None:
            self.mh_network.update_after_iteration()

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


    # -- reference MultiHead:678-901 (validation sweep) and :963-1049 ---------------------------------------------------
    def _perform_validation(self, val_generators, nr_batches, use_head=None, call_for_eval=False):
        """Per-subject validation on every task (reference MultiHead:678-901 minus its plans / dataset / file plumbing):
        `val_generators` maps task -> generator of {'data', 'target', 'keys'} batches.  For each task the matching head is
        assembled (`use_head` when the task has none, :779-783), `nr_batches` batches run without backprop with the
        per-subject tp / fp / fn evaluation (`eval_batch = False`, :703), then Dice / IoU per subject and mask are stored in
        `self.validation_results['epoch_N'][task]` (:963-1049).  Returns that dictionary."""
        self.eval_batch = False
        active = self.task
        if not hasattr(self, "validation_results"):
            self.validation_results = {}
        self.network.eval()
        try:
            for task, gen in val_generators.items():
                head = task if (self.mh_network is None or task in self.mh_network.heads) else use_head
                assert head is not None, ("The task to perform validation/evaluation on is not in the head and no head_name "
                                          "that should be used instead (use_head) is provided.")
                if self.mh_network is not None and head != self.mh_network.active_task:
                    self.network = self.mh_network.assemble_model(head)
                self.task = task
                self.subject_names_raw = []
                with torch.no_grad():
                    for _ in range(nr_batches):
                        batch = next(gen)
                        self.subject_names_raw.append(list(batch['keys']))
                        self.run_iteration(iter([batch]), False, True, no_loss=call_for_eval)
                self.finish_online_evaluation_extended(task)
        finally:
            if self.mh_network is not None and self.mh_network.active_task != active and active in self.mh_network.heads:
                self.network = self.mh_network.assemble_model(active)
            self.task = active
            self.network.train()
            self.eval_batch = True
        return self.validation_results

    def finish_online_evaluation_extended(self, task, unique_subject_names=None):
        """reference MultiHead:963-1049: sum tp / fp / fn over all patches of a subject, then per subject and foreground
        mask IoU = tp / (tp + fp + fn) and Dice = 2 tp / (2 tp + fp + fn) (no smoothing: an absent, unpredicted mask is NaN)"""
        names = np.array(self.subject_names_raw).flatten()
        tp = np.concatenate(self.online_eval_tp, 0).reshape(len(names), -1)
        fp = np.concatenate(self.online_eval_fp, 0).reshape(len(names), -1)
        fn = np.concatenate(self.online_eval_fn, 0).reshape(len(names), -1)
        subjects = list(np.unique(names)) if unique_subject_names is None else list(unique_subject_names)
        store = {}
        for subject in subjects:
            idx = np.where(names == subject)
            i, j, k = tp[idx].sum(0), fp[idx].sum(0), fn[idx].sum(0)
            if np.isnan(i).any():
                continue
            with np.errstate(invalid='ignore', divide='ignore'):
                iou, dc = i / (i + j + k), 2 * i / (2 * i + j + k)
            store[str(subject)] = {'mask_' + str(c + 1): {'IoU': np.float64(iou[c]), 'Dice': np.float64(dc[c])}
                                   for c in range(len(iou))}
        if not hasattr(self, "validation_results"):
            self.validation_results = {}
        self.validation_results.setdefault('epoch_' + str(self.epoch), {})[task] = store
        self.online_eval_tp, self.online_eval_fp, self.online_eval_fn = [], [], []
        self.subject_names_raw = []
        return store

    def predict_preprocessed_data_return_seg_and_softmax(self, data, do_mirroring=True, mirror_axes=(0, 1, 2),
                                                         use_sliding_window=True, step_size=0.5, use_gaussian=True, **_):
        """nnunet nnUNetTrainer.predict_preprocessed_data_return_seg_and_softmax -> SegmentationNetwork.predict_3D
        (what the reference's inference/predict.py:117-401 and evaluation sweep call): tiled sliding-window prediction of
        one preprocessed case, on the CUDA path (b200unet/inference.py).  Returns numpy (segmentation, class probabilities)."""
        from . import inference
        assert use_sliding_window, "only the tiled (sliding-window) prediction is implemented"
        training = self.network.training
        self.network.eval()
        try:
            seg, probs = inference.predict_3D(self.network, data, self.geometry.patch, do_mirroring, tuple(mirror_axes),
                                              step_size, use_gaussian)
        finally:
            self.network.train(training)
        return seg.cpu().numpy(), probs.cpu().numpy()


class nnUNetTrainerSequential(nnUNetTrainerMultiHead):
    """Plain sequential fine-tuning baseline (BASELINE.json config 1): the MultiHead iteration as is."""
    EXTENSION = "sequential"


# ------------------------------------------------------------------------------------------------------------------------
class nnUNetTrainerEWC(nnUNetTrainerMultiHead):
    EXTENSION = "ewc"

    def save_importance(self, path):
        from . import checkpoint
        checkpoint.save_importance(self, path)

    def load_importance(self, path):
        from . import checkpoint
        checkpoint.load_importance(self, path)

    def __init__(self, *a, ewc_lambda=0.4, **kw):
        super().__init__(*a, **kw)
        self.ewc_lambda = ewc_lambda
        self.fisher, self.params = dict(), dict()

    def initialize_loss(self):
        """reference ewc:131-140"""
        self.loss = ds.MultipleOutputLossEWC(self._base_loss(), self.ds_loss_weights, self.ewc_lambda, self.fisher,
                                             self.params, self.network.named_parameters())

    def _net_params(self):
        """Q1: the reference hands the loss a GENERATOR (ewc:247), which the first stored task exhausts"""
        named = self.network.named_parameters()
        return named if self.strict_reference else list(named)

    def _fused_supported(self):
        return type(self.loss) is ds.MultipleOutputLossEWC and hasattr(self.loss.loss, 'cfg')

    def _penalty_names(self):
        return self._matched_param_names()

    def _fused_spec(self):
        tasks = list(self.loss.tasks)
        if self.strict_reference:
            tasks = tasks[:1]
        names = self._penalty_names()
        pen = [(self.ewc_lambda / 2, self.fisher[t], self.params[t], None, names) for t in tasks] if names else []
        key = ("ewc", tuple(tasks), tuple(id(self.fisher[t]) for t in tasks), len(names))
        return key, dict(base='dcce', cfg=self.loss.loss.cfg(list(self.ds_loss_weights)), penalty=pen)

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
        if self._use_fused(do_backprop, no_loss):
            return self._run_iteration_fused(data_generator, run_online_evaluation, detach)
        self._last_penalty = None
        self.loss.update_network_params(self._net_params())
        loss = super().run_iteration(data_generator, do_backprop, run_online_evaluation, False, no_loss)
        if loss is not None:
            if not do_backprop:      # validation iterations: value of the regulariser without touching gradients
                self.loss.update_network_params(self._net_params())
                loss = self.loss._penalty(loss.detach(), self.ewc_lambda / 2)
            elif self._last_penalty is not None:
                loss = loss.detach() + self._last_penalty
            if detach:
                loss = loss.detach().cpu().numpy()
        # reference ewc:247 -- a fresh generator for the next iteration
        self.loss.update_network_params(self._net_params())
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
        self.loss.update_network_params(self._net_params())
        self._steps = {}


class _MaskedEWC(nnUNetTrainerEWC):
    """EWC restricted to name-matched parameters (reference ewc_vit / ewc_ln / ewc_unet: `EWCLoss(..., True, MATCH,
    MATCH_TRUE)`), SURVEY 8(f) rank 3.  After every iteration the reference hands the loss a LIST of the selected
    parameters (ewc_vit:68-69), so -- unlike plain EWC (Q1) -- every stored task is penalised; `after_train` drops the
    Fisher / parameter entries outside the mask (ewc_vit:79-88)."""
    MATCH, MATCH_TRUE = ['ViT'], True

    def initialize_loss(self):
        self.loss = ds.MultipleOutputLossEWC(self._base_loss(), self.ds_loss_weights, self.ewc_lambda, self.fisher, self.params,
                                             self._net_params(), True, list(self.MATCH), self.MATCH_TRUE)

    def _selected(self, name):
        return ds._match(name, True, self.MATCH, self.MATCH_TRUE)

    def _net_params(self):
        return [(n, p) for n, p in self.network.named_parameters() if self._selected(n)]

    def _fused_spec(self):
        key, spec = super()._fused_spec()
        names = self._penalty_names()
        tasks = list(self.loss.tasks)            # a list is handed over: no generator exhaustion
        spec['penalty'] = [(self.loss.ewc_lambda / 2, self.fisher[t], self.params[t], None, names) for t in tasks] if names else []
        return ("ewc-masked", tuple(tasks), tuple(id(self.fisher[t]) for t in tasks), len(names), float(self.loss.ewc_lambda)), spec

    def _penalty_after_backward(self):
        return self.loss.penalty_into_grads(self.loss.ewc_lambda / 2)

    def after_train(self, data_generator, num_batches=1):
        super().after_train(data_generator, num_batches)
        for d in (self.fisher, self.params):
            for task in list(d.keys()):
                for key in list(d[task].keys()):
                    if not self._selected(key):
                        del d[task][key]
        self.loss.update_ewc_params(self.fisher, self.params)
        self.loss.update_network_params(self._net_params())


class nnUNetTrainerEWCViT(_MaskedEWC):
    """reference ewc_vit/nnUNetTrainerEWCViT.py:33-88 -- EWC on the ViT parameters only"""
    MATCH, MATCH_TRUE = ['ViT'], True


class nnUNetTrainerEWCLN(_MaskedEWC):
    """reference ewc_ln/nnUNetTrainerEWCLN.py:33-98 -- EWC on the ViT's LayerNorm parameters only"""
    MATCH, MATCH_TRUE = ['ViT', 'norm'], True


class nnUNetTrainerEWCUNet(_MaskedEWC):
    """reference ewc_unet/nnUNetTrainerEWCUNet.py:33-90 -- EWC on everything except the ViT"""
    MATCH, MATCH_TRUE = ['ViT'], False


class _FreezeAfterFirstTask(nnUNetTrainerMultiHead):
    """reference frozen_vit / frozen_unet / frozen_nonln `run_training` (:28-62): when the SECOND task starts, the selected
    parameters are frozen for good (running model, body and first head share the Parameter objects here)."""

    def __init__(self, *a, **kw):
        super().__init__(*a, **kw)
        self.frozen = False

    def _freeze(self, name):
        raise NotImplementedError

    def start_task(self, task):
        if not self.frozen and len(self.mh_network.heads) == 1 and task not in self.mh_network.heads:
            for mod in (self.network, self.mh_network.body, self.mh_network.heads[list(self.mh_network.heads.keys())[0]]):
                for name, param in mod.named_parameters():
                    if self._freeze(name):
                        param.requires_grad = False
            self.frozen = True
        super().start_task(task)


class nnUNetTrainerFrozenViT(_FreezeAfterFirstTask):
    def _freeze(self, name):
        return 'ViT' in name                                    # frozen_vit:36-39


class nnUNetTrainerFrozenUNet(_FreezeAfterFirstTask):
    def _freeze(self, name):
        return 'ViT' not in name                                # frozen_unet:36-40


class nnUNetTrainerFrozenNonLN(_FreezeAfterFirstTask):
    def _freeze(self, name):
        return 'norm' not in name or 'ViT' not in name          # frozen_nonln:36-43: everything but the ViT's LayerNorms


class nnUNetTrainerFrozenBody(nnUNetTrainerMultiHead):
    """reference frozen_body_seq/nnUNetTrainerFrozenUNet.py:169-229: from the second task on the shared body is frozen
    (`assemble_model(task, freeze_body=True)`), only the task's head -- whose parameters are re-enabled explicitly (:203-204) --
    is trained; the first task trains everything."""
    EXTENSION = "frozen_body_seq"

    def start_task(self, task):
        if task not in self.mh_network.heads:
            self.mh_network.add_new_task(task, use_init=not self.transfer_heads)
        for _, param in self.mh_network.heads[task].named_parameters():
            param.requires_grad = True
        self.network = self.mh_network.assemble_model(task, freeze_body=len(self.mh_network.heads) > 1)
        self.task = task
        self.already_trained_on[str(self.fold)]["start_training_on"] = task
        self._steps = {}
        self.initialize_optimizer_and_scheduler()            # :226


class nnUNetTrainerRehearsal(nnUNetTrainerMultiHead):
    """reference rehearsal/nnUNetTrainerRehearsal.py:65-173: the training set of a task is fused with a seeded sample
    (`samples_in_perc`) of the training cases of every task already in the heads; the iteration itself is the plain MultiHead one.
    Here the fused case list feeds the GPU patch pipeline (b200unet/augment.py) instead of nnunet's DataLoader3D."""
    EXTENSION = "rehearsal"

    def __init__(self, *a, samples_in_perc=0.25, rehearsal_seed=3299, **kw):
        super().__init__(*a, **kw)
        assert 0 < samples_in_perc <= 1, "samples_in_perc should be between 0 and 1: (0, 1]."       # rehearsal:30-31
        self.samples, self.rehearsal_seed = samples_in_perc, rehearsal_seed

    def fuse_datasets(self, cases_current, cases_by_previous_task):
        """cases_*: {case key: case dict}.  Returns the fused training dict (:80-124): `random.seed(seed)`, then per previous
        task -- in head order -- `random.sample(items, round(len * samples))`; later tasks overwrite equal keys"""
        import random
        random.seed(self.rehearsal_seed)
        fused = dict(cases_current)
        done = [t for t in (self.mh_network.heads.keys() if self.mh_network is not None else []) if t in cases_by_previous_task]
        for task in done:
            items = list(cases_by_previous_task[task].items())
            fused.update(random.sample(items, round(len(items) * self.samples)))
        random.seed()
        return fused

    def get_basic_generators(self, cases_current, cases_by_previous_task, cases_val, ds_strides=None, **pipeline_kw):
        """(training pipeline over the fused cases, validation pipeline of the current task), as :151-171 builds the DataLoader3D pair"""
        from . import augment
        if ds_strides is None:
            ds_strides, cum = [(1, 1, 1)], [1, 1, 1]
            for k in self.geometry.pool[:-1]:
                cum = [a * b for a, b in zip(cum, k)]
                ds_strides.append(tuple(cum))
        fused = self.fuse_datasets(cases_current, cases_by_previous_task)
        mk = lambda cases, train: augment.GPUPatchPipeline([dict(c, key=k) for k, c in cases.items()], self.geometry.patch,
                                                           self.geometry.batch, ds_strides, device=self.device, train=train, **pipeline_kw)
        return mk(fused, True), mk(cases_val, False)


class nnUNetTrainerFrozEWC(nnUNetTrainerEWCViT):
    """reference froz_ewc/nnUNetTrainerFrozEWC.py:81-161: the ViT is frozen on every second task and EWC-regularised on the
    others; `adaptive` scales the EWC weight by e^(-1/3) while the ViT is frozen (:107,:117)."""

    def __init__(self, *a, adaptive=False, **kw):
        super().__init__(*a, **kw)
        self.adaptive = adaptive

    def freeze_ViT(self, freeze):
        for mod in (self.network, self.mh_network.body, self.mh_network.heads[list(self.mh_network.heads.keys())[0]]):
            for name, param in mod.named_parameters():
                if 'ViT' in name:
                    param.requires_grad = not freeze

    def start_task(self, task):
        n_heads, known = len(self.mh_network.heads), task in self.mh_network.heads
        freeze = (not known) if n_heads % 2 == 1 else known     # froz_ewc:92-130
        self.freeze_ViT(freeze)
        if self.adaptive:
            self.loss.ewc_lambda = self.ewc_lambda * math.exp(-1 / 3) if freeze else self.ewc_lambda
        super().start_task(task)


# ------------------------------------------------------------------------------------------------------------------------
class nnUNetTrainerRW(nnUNetTrainerMultiHead):
    EXTENSION = "rw"

    def save_importance(self, path):
        from . import checkpoint
        checkpoint.save_importance(self, path)

    def load_importance(self, path):
        from . import checkpoint
        checkpoint.load_importance(self, path)

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
        """reference rw:160-168 (+ the head bookkeeping of MultiHead.run_training :541-564)"""
        super().start_task(task)
        self.params[task] = dict()
        self.fisher[task] = {n: torch.zeros_like(p, requires_grad=False) for n, p in self.network.named_parameters() if p.requires_grad}
        self.scores[task] = {n: torch.zeros_like(p, requires_grad=False) for n, p in self.network.named_parameters() if p.requires_grad}
        self.loss.update_rw_params(self.fisher, self.params, self.scores)
        # Q2: the reference hands the loss a generator ONCE and never refreshes it: the penalty is evaluated on the first
        # iteration after the stored tasks become non-empty and is silently zero afterwards (strict_reference); the
        # documented math (deep_supervision.py:115-132) penalises every iteration (strict_reference=False)
        self._rw_params_fresh = True
        self.loss.update_network_params(self.network.named_parameters() if self.strict_reference
                                        else list(self.network.named_parameters()))

    def _fused_supported(self):
        return type(self.loss) is ds.MultipleOutputLossRW and hasattr(self.loss.loss, 'cfg')

    def _fused_spec(self):
        tasks = list(self.loss.tasks)
        if self.strict_reference:
            tasks = tasks[:1] if getattr(self, "_rw_params_fresh", False) else []
        names = self._matched_param_names()
        pen = [(self.rw_lambda, self.fisher[t], self.params[t], self.scores[t], names) for t in tasks] if names else []
        key = ("rw", tuple(tasks), tuple(id(self.fisher[t]) for t in tasks), tuple(id(self.scores[t]) for t in tasks))
        return key, dict(base='dcce', cfg=self.loss.loss.cfg(list(self.ds_loss_weights)), penalty=pen)

    def _fused_post(self, step):
        if self.strict_reference and getattr(self, "_rw_params_fresh", False) and self.loss.tasks:
            self._rw_params_fresh = False
            self.loss.update_network_params(iter(()))         # the generator is exhausted now (Q2)
        self._update_f_s_values()

    def _forward_loss(self, data, target):
        output = self.network(data)
        return output, ds.MultipleOutputLossEWC.forward(self.loss, output, target, reg=False)

    def _backward(self, l):
        l.backward()
        self._sync_gradients()
        self._last_penalty = self.loss.penalty_into_grads(self.rw_lambda, self.loss.parameter_importance)

    def run_iteration(self, data_generator, do_backprop=True, run_online_evaluation=False, detach=True, no_loss=False):
        if self._use_fused(do_backprop, no_loss):
            return self._run_iteration_fused(data_generator, run_online_evaluation, detach)
        self._last_penalty = None
        loss = super().run_iteration(data_generator, do_backprop, run_online_evaluation, False, no_loss)
        if do_backprop and self.loss.tasks:
            self._rw_params_fresh = False
        if loss is not None:
            if not do_backprop:      # validation iterations: the reference's forward always adds the penalty (deep_supervision.py:109-135)
                loss = self.loss._penalty(loss.detach(), self.rw_lambda, self.loss.parameter_importance)
            elif self._last_penalty is not None:
                loss = loss.detach() + self._last_penalty
        if do_backprop:
            self._update_f_s_values()
        else:
            # rw:224 calls _update_f_s_values on validation iterations too, where its effect depends on the torch version
            # (zero_grad leaves zeros under torch 1.9 -> F decays, None under torch >= 2 -> prev_param = {} and the next
            # update raises KeyError); here validation iterations only advance the schedule counter
            self.count += 1
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

    def _fused_supported(self):
        return self.network_old is None or type(self.network_old) is Generic_UNet

    def _fused_spec(self):
        cfg = self.loss_base.loss.cfg(list(self.ds_loss_weights))
        if self.network_old is None:
            return ("mib0",), dict(base='dcce', cfg=cfg)
        return ("mib", id(self.network_old)), dict(base='ce255', cfg=cfg, teacher=self.network_old, mib=(self.alpha, self.lkd))

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

    def _fused_supported(self):
        return self.network_old is None or type(self.network_old) is Generic_UNet

    def _fused_spec(self):
        cfg = self.loss_base.loss.cfg(list(self.ds_loss_weights))
        if self.network_old is None:
            return ("pod0",), dict(base='dcce', cfg=cfg)
        return ("pod", id(self.network_old)), dict(base='dcce', cfg=cfg, teacher=self.network_old,
                                                   pod=(self.pod_lambda, self.scales))

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

    @staticmethod
    def _median_thresholds(hist, nb_bins=100, base_threshold=0.001):
        """median entropy per pseudo label from the histogram (arthurdouillard/CVPR2021_PLOP train.py find_median, which
        plop:149-173 follows), floored at base_threshold"""
        out = []
        for c in range(hist.shape[0]):
            total, med = float(hist[c].sum()), 0.0
            if total > 0:
                half, running = total / 2, 0.0
                for bin_index in range(nb_bins):
                    lower = bin_index / nb_bins
                    if half >= running and half <= running + float(hist[c, bin_index]):
                        break
                    running += float(hist[c, bin_index])
                med = lower + ((half - running) / float(hist[c, bin_index])) * (1 / nb_bins)
            out.append(max(med, base_threshold))
        return out

    def extract_max_entropy_and_thresholds(self, data_generator, num_batches):
        """reference plop:113-182: `num_batches` batches through the OLD model; per deep-supervision level the median of
        entropy / max_entropy over the background voxels, per pseudo label, becomes the pseudo-labelling threshold.
        strict_reference (Q15): the reference compares the LIST of deep-supervision targets with 0 (`labels == 0` is a Python
        False), so nothing is selected, the histograms stay empty and every threshold is base_threshold = 0.001 -- the batches
        are still drawn so the data stream stays aligned.  strict_reference=False runs the documented algorithm: histograms
        by b2_plop_entropy_hist (ONE table accumulated over batches and levels, as the reference's shared `histograms`),
        thresholds[level] taken after that level of the LAST batch."""
        import ctypes as C
        from . import _lib
        C_, nb = self.geometry.num_classes, 100
        self.max_entropy = torch.log(torch.tensor(C_).float()).item()
        levels = self.geometry.num_pool
        if self.strict_reference or self.network_old is None:
            for _ in range(num_batches):
                next(data_generator)
            self.thresholds = {i: torch.full((C_,), 1e-3, device=self.device) for i in range(levels)}
            self._steps = {}
            return self.thresholds
        lib = _lib.load()
        hist = torch.zeros((C_, nb), dtype=torch.int64, device=self.device)
        st = lambda: C.c_void_p(torch.cuda.current_stream(self.device).cuda_stream)
        self.network_old.eval()
        with torch.no_grad():
            for k in range(num_batches):
                data, target = self._next_batch(data_generator)
                outs = self.network_old(data)
                for idx in range(levels):
                    o, t = outs[idx].contiguous().float(), target[idx].contiguous().float()
                    V = o[0, 0].numel()
                    _lib.check(lib.b2_plop_entropy_hist(C.c_void_p(o.data_ptr()), C.c_void_p(t.data_ptr()), int(o.shape[0]), C_, V,
                                                        float(self.max_entropy), nb, C.c_void_p(hist.data_ptr()), st()))
                    if k == num_batches - 1:
                        self.thresholds[idx] = torch.tensor(self._median_thresholds(hist.cpu().numpy(), nb), dtype=torch.float32,
                                                            device=self.device)
        self._plop_histograms = hist
        self._steps = {}            # captured step programs hold the previous thresholds
        return self.thresholds

    def _fused_spec(self):
        cfg = self.loss_base.loss.cfg(list(self.ds_loss_weights))
        if self.network_old is None:
            return ("plop0",), dict(base='dcce', cfg=cfg)
        thr = [self.thresholds[i].to(self.device).contiguous().float() for i in range(self.geometry.num_pool)]
        return (("plop", id(self.network_old), float(self.max_entropy)),
                dict(base='plop', cfg=cfg, teacher=self.network_old, pod=(self.pod_lambda, self.scales),
                     plop=(thr, self.max_entropy)))

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

    def __init__(self, *a, lwf_temperature=2.0, **kw):
        super().__init__(*a, **kw)
        self.lwf_temperature = lwf_temperature
        self.heads = dict()          # finished task -> state_dict snapshot of its head (the split module)
        self.batch_idx = 0           # running index into the stored target logits (lwf:362)
        self.target_logits = dict()  # task -> list of stored full-res logits (lwf:247-251)

    def initialize_loss(self):
        self.loss = ds.MultipleOutputLossLWF(self._base_loss(), self.ds_loss_weights, list(), list(), self.lwf_temperature)

    def _head_module(self, net=None):
        m = self.network if net is None else net
        for part in self.split.split('.'):
            m = getattr(m, part)
        return m

    def finish_task(self):
        """end of a task: its head is frozen (MultiHead_Module keeps it under ``mh_network.heads[task]``; the snapshot below
        is what the distillation reads)"""
        self.heads[self.task] = {k: v.detach().clone() for k, v in self._head_module().state_dict().items()}
        self._steps = {}

    def _fused_supported(self):
        return (type(self.loss) is ds.MultipleOutputLossLWF and hasattr(self.loss.loss, 'cfg') and
                (not self.heads or self._shared_body()))

    def _fused_spec(self):
        cfg = self.loss.loss.cfg(list(self.ds_loss_weights))
        if not self.heads:
            return ("lwf0",), dict(base='dcce', cfg=cfg)
        P = self.geometry.num_pool
        heads = [(t, self.heads[t]["%d.weight" % (P - 1)]) for t in self.heads]
        return ("lwf", tuple(self.heads), tuple(id(self.heads[t]) for t in self.heads)), \
            dict(base='dcce', cfg=cfg, lwf=(heads, self.lwf_temperature))

    def _fused_pre(self, step):
        if step.lwf is not None:
            idx = self._cur_batch_idx
            for h in step.lwf['heads']:
                stored = self.target_logits[h['task']]
                h['target'].copy_(stored[idx % len(stored)], non_blocking=True)

    def _shared_body(self):
        """the documented split (`-s seg_outputs`, run_training.py:103-107): a stored head is only the 1x1x1 output
        convolutions, so every old head can be evaluated on the body activations of ONE forward"""
        return self.split == "seg_outputs" and hasattr(self.network, "head_logits")

    def _old_head_logits(self, task):
        """full-resolution logits of stored head `task` on the activations of the last forward"""
        P = self.geometry.num_pool
        return self.network.head_logits(self.heads[task]["%d.weight" % (P - 1)], level=0)

    def _forward_with_head(self, data, task):
        """assemble_model(task) + eval forward + first output (lwf:315-346), without touching the training state"""
        if self._shared_body():
            with torch.no_grad():
                self.network(data)
            return self._old_head_logits(task)
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
        self.target_logits = {task: [] for task in self.heads}
        for d in data_batches:
            data = to_cuda(maybe_to_torch(d), gpu_id=self.device.index)
            if self._shared_body():
                with torch.no_grad():
                    self.network(data)                   # one body forward per batch, every stored head on top of it
                for task in self.heads:
                    self.target_logits[task].append(self._old_head_logits(task))
            else:
                for task in self.heads:
                    self.target_logits[task].append(self._forward_with_head(data, task))

    def _forward_loss(self, data, target):
        if not self.heads:
            output = self.network(data)
            return output, self.loss(output, target)
        # the old heads' predictions of THIS batch (Q16: the intended "same batch"); with the documented split they are
        # evaluated on the body activations the training forward just produced: same parameters, and eval == train for
        # InstanceNorm without running stats / dropout p = 0
        if self._shared_body():
            output = self.network(data)
            preds = [self._old_head_logits(task) for task in self.heads]
        else:
            preds = [self._forward_with_head(data, task) for task in self.heads]
            output = self.network(data)
        idx = self._cur_batch_idx
        targets = [self.target_logits[task][idx % len(self.target_logits[task])] for task in self.heads]
        self.loss.update_logits(preds, targets)
        return output, self.loss(output, target)

    def run_iteration(self, data_generator, do_backprop=True, run_online_evaluation=False, detach=True, no_loss=False,
                      batch_idx=None):
        if batch_idx is None:
            batch_idx = self.batch_idx
            if do_backprop:
                self.batch_idx += 1
        self._cur_batch_idx = batch_idx
        return super().run_iteration(data_generator, do_backprop, run_online_evaluation, detach, no_loss)
