"""The training iteration as ONE enqueue program over the C ABI (include/b2unet.h) -- no autograd graph, no per-step
allocations, no per-step host tables -- optionally captured in a CUDA graph and replayed.

What it replaces: the body of ``nnUNetTrainerMultiHead.run_iteration`` (reference .../multihead/nnUNetTrainerMultiHead.py:
617-641: zero_grad, forward, loss, backward, clip_grad_norm_, optimizer.step) and the loss compositions of
reference loss_functions/deep_supervision.py (EWC :58-83, RW :109-135, LwF :201-214, MiB :401-416, PLOP :247-285,
POD :361-380) for the trainers in b200unet/trainers.py.  The public ``Generic_UNet`` / ``MultipleOutputLoss*`` classes keep
their autograd interface for drop-in use; the trainers take this path because at a ~5 ms GPU step the Python / autograd /
allocator work of the generic path (3 ms of host time, ~250 launches) is of the same order as the step itself.

Sequence enqueued by ``FusedStep.enqueue`` (all on the current stream unless noted):
  [teacher forward (side stream, own plan)]  student forward  ->  per deep-supervision level: Dice+CE | CE(ignore 255)
  | PLOP pseudo-label CE (value + dlogits; levels >= 1 on side streams) [+ MiB KD into the same dlogits]
  [+ POD / LwF value-only terms]  ->  backward with gradient buckets  ->  [bucketed NCCL all-reduce(AVG) on a communication
  stream, overlapped with the remaining backward]  ->  [EWC / RW penalty: value + gradient, one launch per stored task]
  ->  clip(12) + SGD-Nesterov (lr read from device memory)  ->  loss terms summed in a fixed order.
``param.grad`` of every parameter is a persistent view of one flat arena (overwritten every step), so RW / EWC / user
code read gradients exactly where the reference leaves them.
"""
import ctypes as C

import torch

from . import _lib


def _st(dev):
    return C.c_void_p(torch.cuda.current_stream(dev).cuda_stream)


def _vp(t):
    return None if t is None else C.c_void_p(t.data_ptr())


class DeviceTable:
    """A multi-tensor table built once on the host (b2_mt_blob_build) and kept in device memory."""

    def __init__(self, kind, entries, n, dev, keep=()):
        lib = _lib.load()
        nbytes = int(lib.b2_mt_blob_bytes(kind, n))
        host = (C.c_char * nbytes)()
        nblocks = C.c_int32()
        _lib.check(lib.b2_mt_blob_build(kind, entries, n, host, C.byref(nblocks)))
        self.blob = torch.frombuffer(host, dtype=torch.uint8).clone().to(dev)
        self.n, self.nblocks = n, int(nblocks.value)
        self.part = torch.empty(int(lib.b2_mt_part_bytes(self.nblocks)), dtype=torch.uint8, device=dev)
        self.keep = keep          # tensors the entries point to


def bucket_bounds(names, numels, target_bytes):
    """Gradient buckets for the overlapped all-reduce: [(first_param, end_param)] in the order backward completes them -- suffixes of
    the plan's parameter order (backward finishes layers in DESCENDING parameter index), cut only at layer starts (a conv / tconv
    / head weight), each at least `target_bytes` of fp32 gradients except the last, which takes what is left down to parameter 0."""
    starts = [i for i, n in enumerate(names) if n.endswith("conv.weight") or
              (n.endswith(".weight") and (n.startswith("tu.") or n.startswith("seg_outputs.")))]
    bounds, cur_hi, acc, prev = [], len(names), 0, len(names)
    for s in reversed(starts):               # greedy from the end of the arena
        acc += sum(numels[s:prev]) * 4
        prev = s
        if acc >= target_bytes or s == 0:
            bounds.append((s, cur_hi))
            cur_hi, acc = s, 0
    if not bounds or bounds[-1][0] != 0:
        bounds.append((0, cur_hi))
    return bounds


class FusedStep:
    def __init__(self, trainer, data, targets, spec):
        self.tr, self.spec = trainer, spec
        self.lib = _lib.load()
        net = trainer.network
        dev = data.device
        self.dev = dev
        self.plan = net._get_plan(data)
        plan = self.plan
        self.params = net._ordered_params(plan)
        self.param_ptrs = _ptr_array(self.params)
        self.data = torch.empty(tuple(data.shape), dtype=torch.float32, device=dev)
        self.targets = [torch.empty(tuple(t.shape), dtype=torch.float32, device=dev) for t in targets]
        self.logits = [torch.empty(s, dtype=torch.float32, device=dev) for s in plan.out_shapes]
        self.logit_ptrs = _ptr_array(self.logits)
        P = len(self.logits)
        w = [float(x) for x in trainer.ds_loss_weights]
        self.level_on = [i == 0 or w[i] != 0 for i in range(P)]
        self.weights = w
        self.dl = [torch.zeros_like(self.logits[i]) if self.level_on[i] else None for i in range(P)]
        self.dl_ptrs = _ptr_array(self.dl)
        self.loss_scr = [torch.empty(int(self.lib.b2_dsloss_scratch_bytes(int(l.shape[0]), int(l.shape[1]), l[0, 0].numel())) +
                                     int(self.lib.b2_kd_scratch_bytes(int(l.shape[0]), int(l.shape[1]), l[0, 0].numel())),
                                     dtype=torch.uint8, device=dev) if self.level_on[i] else None
                         for i, l in enumerate(self.logits)]
        # gradient arena; p.grad = persistent views
        total = sum(plan.param_numel)
        self.flat = torch.zeros(total, dtype=torch.float32, device=dev)
        self.grads, o = [], 0
        self.offsets = []
        for n_, s in zip(plan.param_numel, plan.param_shapes):
            self.grads.append(self.flat[o:o + n_].view(s))
            self.offsets.append(o)
            o += n_
        self.grad_ptrs = _ptr_array(self.grads)
        self.has = (C.c_int32 * len(self.grads))()
        # teacher (MiB / PLOP / POD)
        self.teacher = spec.get('teacher')
        if self.teacher is not None:
            self.tplan = self.teacher._get_plan(data)
            self.tparams = self.teacher._ordered_params(self.tplan)
            self.tparam_ptrs = _ptr_array(self.tparams)
            self.tlogits = [torch.empty(s, dtype=torch.float32, device=dev) for s in self.tplan.out_shapes]
            self.tlogit_ptrs = _ptr_array(self.tlogits)
            self.tstream = torch.cuda.Stream(dev)
        # loss term slots
        self.n_terms = 0
        self.slot_level = [self._slot() if on else None for on in self.level_on]
        self.pod = None
        if spec.get('pod'):
            self._build_pod()
        self.lwf = None
        if spec.get('lwf'):
            self._build_lwf()
        self.pen = []
        for coef, fisher, stars, importance, names in spec.get('penalty', []):
            self.pen.append(self._build_penalty(coef, fisher, stars, importance, names))
        self.terms = torch.zeros(max(self.n_terms, 1), dtype=torch.float32, device=dev)
        self.total = torch.zeros((), dtype=torch.float32, device=dev)
        self.level_streams = [None] + [torch.cuda.Stream(dev) for _ in range(1, P)]
        # optimiser table: parameters that end up with a gradient = data-term gradients (known after the first backward)
        # + everything the penalty touches
        self.sgd = None
        self.hyper = torch.zeros(4, dtype=torch.float32, device=dev)
        self._hyper_host = None
        self.norm = torch.zeros(1, dtype=torch.float32, device=dev)
        # data parallel
        self.ddp = trainer.ddp
        self.buckets, self.bucket_arr = [], None
        if self.ddp is not None:
            self._build_buckets()
        self.graph = None
        self.steps_run = 0

    # ------------------------------------------------------------------------------------------------------------
    def _slot(self, n=1):
        s = self.n_terms
        self.n_terms += n
        return s

    def _build_pod(self):
        """static (student, teacher) raw conv-output pairs in hook order (reference plop:330-353 hooks every module whose
        type string contains 'conv.Conv': the 3x3x3 convs, the transposed convs and the 1x1x1 heads)"""
        from .deep_supervision import _act_view
        lam, scales = self.spec['pod']
        net, P = self.tr.network, self.tr.network.num_pool
        n_enc = 2 * (P + 1)
        pairs = []
        for i in range(len(self.plan.conv_names)):
            a, _ = self.plan.conv_output(i)
            b, _ = self.tplan.conv_output(i)
            pairs.append((a, b))
            if i >= n_enc and (i - n_enc) % 3 == 2:
                u = (i - n_enc) // 3
                pairs.append((self.logits[P - 1 - u], self.tlogits[P - 1 - u]))
        L = len(pairs)
        self.pod = dict(lam=lam, scales=int(scales), L=L, slot=self._slot(), views=[], keep=[], refresh=[])
        self.pod_vals = torch.zeros(L, dtype=torch.float32, device=self.dev)
        # Q5: `dist = (dist + lam * pod_l) / L` inside the loop  ==  sum_l lam * pod_l / L^(L-l)
        self.pod_coef = torch.tensor([lam / float(L) ** (L - l) for l in range(L)], dtype=torch.float64, device=self.dev)

        def static_view(t):
            """b2_act_view of a persistent tensor; NCDHW logits get a persistent channels-last copy refreshed every step"""
            st = t.stride()
            if not (t.dim() == 5 and st[1] == 1):
                cl = torch.empty_like(t, memory_format=torch.channels_last_3d)
                self.pod['refresh'].append((t, cl))
                t = cl
            v, k = _act_view(t)
            assert k.data_ptr() == t.data_ptr(), "POD operand was copied: the view would go stale"
            return v, k
        for a, b in pairs:
            va, ka = static_view(a)
            vb, kb = static_view(b)
            scr = torch.empty(int(self.lib.b2_pod_scratch_bytes(C.byref(va), int(scales))), dtype=torch.uint8, device=self.dev)
            self.pod['views'].append((va, vb, scr))
            self.pod['keep'].append((ka, kb))

    def _build_lwf(self):
        heads, T = self.spec['lwf']            # [(task, full-resolution head weight)], temperature
        l0 = self.logits[0]
        self.lwf = dict(T=float(T), heads=[])
        for task, wgt in heads:
            self.lwf['heads'].append(dict(task=task, w=wgt.detach().to(self.dev, torch.float32).contiguous(),
                                          pred=torch.empty_like(l0), target=torch.empty_like(l0), slot=self._slot()))
        B, Cc, V = int(l0.shape[0]), int(l0.shape[1]), l0[0, 0].numel()
        self.lwf_scr = torch.empty(int(self.lib.b2_kd_scratch_bytes(B, Cc, V)), dtype=torch.uint8, device=self.dev)

    def _build_penalty(self, coef, fisher, stars, importance, names):
        table = (_lib.PenEntry * len(names))()
        keep = []
        for i, n in enumerate(names):
            k = self.plan.param_names.index(n)
            p = self.params[k]
            conv = lambda t: t.detach().to(p.device, torch.float32).contiguous()   # no copy when already in place
            f, st = conv(fisher[n]), conv(stars[n])
            if f.numel() != p.numel():          # grad None -> tensor([1]) (ewc:300-301): broadcast
                f = f.expand_as(p).contiguous()
            im = None if importance is None else conv(importance[n])
            for t in (f, st) + (() if im is None else (im,)):
                if t.numel() != p.numel():
                    raise ValueError("EWC/RW state of %s has %d elements, the parameter %d" % (n, t.numel(), p.numel()))
            keep.append((f, st, im))
            table[i].theta, table[i].theta_star, table[i].fisher = p.data_ptr(), st.data_ptr(), f.data_ptr()
            table[i].importance = None if im is None else im.data_ptr()
            table[i].grad, table[i].numel = self.grads[k].data_ptr(), p.numel()
        return dict(coef=float(coef), table=DeviceTable(_lib.MT_PEN, table, len(names), self.dev, keep), slot=self._slot(),
                    names=set(names))

    def _build_buckets(self, target_bytes=None):
        """suffixes of the arena in the order backward completes them (decoder levels from full resolution down, then the
        encoder from the bottleneck up); boundaries only at layer starts.  Bucket size: env B2_BUCKET_MB (default 48)."""
        import os
        if target_bytes is None:
            target_bytes = int(float(os.environ.get("B2_BUCKET_MB", "48")) * (1 << 20))
        names = self.plan.param_names
        bounds = bucket_bounds(names, self.plan.param_numel, target_bytes)
        self.comm = torch.cuda.Stream(self.dev)
        arr = (_lib.GradBucket * len(bounds))()
        for k, (lo, hi) in enumerate(bounds):
            em, es = torch.cuda.Event(), torch.cuda.Event()
            e0 = self.offsets[lo]
            e1 = self.offsets[hi] if hi < len(names) else self.flat.numel()
            self.buckets.append(dict(lo=lo, hi=hi, view=self.flat[e0:e1], em=em, es=es, done=torch.cuda.Event()))
        self.bucket_arr = arr

    def _bucket_struct(self):
        """cudaEvent_t handles exist only after a first record: create them lazily by recording once"""
        for k, b in enumerate(self.buckets):
            for e in (b['em'], b['es']):
                if e.cuda_event == 0:
                    e.record(torch.cuda.current_stream(self.dev))
            self.bucket_arr[k].first_param = b['lo']
            self.bucket_arr[k].event_main = b['em'].cuda_event
            self.bucket_arr[k].event_side = b['es'].cuda_event
        return self.bucket_arr

    def _build_sgd(self):
        opt = self.tr.optimizer
        group = opt.param_groups[0]
        pen_names = set()
        for pe in self.pen:
            pen_names |= pe['names']
        sel = [k for k in range(len(self.params)) if (self.has[k] or self.plan.param_names[k] in pen_names) and self.params[k].requires_grad]
        table = (_lib.SgdEntry * len(sel))()
        keep = []
        for i, k in enumerate(sel):
            p = self.params[k]
            st = opt.state[p]
            if 'momentum_buffer' not in st:
                st['momentum_buffer'] = torch.zeros_like(p)
            table[i].theta, table[i].grad = p.data_ptr(), self.grads[k].data_ptr()
            table[i].momentum, table[i].numel = st['momentum_buffer'].data_ptr(), p.numel()
            keep.append(st['momentum_buffer'])
        self.sgd = DeviceTable(_lib.MT_SGD, table, len(sel), self.dev, keep)
        self.sgd_sel = sel
        self.nesterov = int(group['nesterov'])
        # param.grad = persistent arena views (None for parameters without gradient, like autograd leaves them)
        for k, p in enumerate(self.params):
            p.grad = self.grads[k] if k in set(sel) else None

    def set_hyper(self, max_norm=12.0):
        g = self.tr.optimizer.param_groups[0]
        h = (float(g['lr']), float(g['momentum']), float(g['weight_decay']), float(max_norm))
        if h != self._hyper_host:
            self.hyper.copy_(torch.tensor(h, dtype=torch.float32))
            self._hyper_host = h

    # ------------------------------------------------------------------------------------------------------------
    def enqueue(self):
        self.enqueue_forward()
        self.enqueue_rest()

    def enqueue_forward(self):
        """part A: needs only the input patch (the targets may still be in flight on the copy stream)"""
        lib, dev, plan = self.lib, self.dev, self.plan
        cur = torch.cuda.current_stream(dev)
        ws = C.c_void_p(plan.workspace.data_ptr())
        self.terms.zero_()
        plan.generation += 1
        fork = torch.cuda.Event()
        fork.record(cur)
        # teacher forward next to the student's (frozen copy, own plan and workspace)
        t_done = None
        if self.teacher is not None:
            self.tstream.wait_event(fork)
            with torch.cuda.stream(self.tstream):
                self.tplan.generation += 1
                _lib.check(lib.b2_unet_forward(self.tplan.handle, self.tparam_ptrs, _vp(self.data),
                                               C.c_void_p(self.tplan.workspace.data_ptr()), self.tlogit_ptrs, 0, _st(dev)))
                t_done = torch.cuda.Event()
                t_done.record(self.tstream)
        _lib.check(lib.b2_unet_forward(plan.handle, self.param_ptrs, _vp(self.data), ws, self.logit_ptrs, 1, _st(dev)))
        if t_done is not None:
            cur.wait_event(t_done)

    def enqueue_rest(self):
        """part B: loss terms, backward (+ all-reduce), penalty, optimiser"""
        lib, dev, plan = self.lib, self.dev, self.plan
        cur = torch.cuda.current_stream(dev)
        ws = C.c_void_p(plan.workspace.data_ptr())
        # ---- loss terms --------------------------------------------------------------------------------------------
        mid = torch.cuda.Event()
        mid.record(cur)
        joins = []
        base = self.spec.get('base', 'dcce')
        for i, x in enumerate(self.logits):
            if not self.level_on[i]:
                continue
            side = i > 0
            stream = self.level_streams[i] if side else cur
            if side:
                stream.wait_event(mid)
            with torch.cuda.stream(stream):
                self._level_loss(i, x, base)
                if side:
                    ev = torch.cuda.Event()
                    ev.record(stream)
                    joins.append(ev)
        if self.lwf is not None:
            for h in self.lwf['heads']:
                _lib.check(lib.b2_unet_head_forward(plan.handle, ws, 0, _vp(h['w']), _vp(h['pred']), _st(dev)))
                l0 = self.logits[0]
                _lib.check(lib.b2_kd_lwf(_vp(h['pred']), _vp(h['target']), int(l0.shape[0]), int(l0.shape[1]), l0[0, 0].numel(),
                                         self.lwf['T'], _vp(self.terms[h['slot']:]), _vp(self.lwf_scr), _st(dev)))
        if self.pod is not None:
            for src, dst in self.pod['refresh']:
                dst.copy_(src)
            for l, (va, vb, scr) in enumerate(self.pod['views']):
                _lib.check(lib.b2_pod_local(C.byref(va), C.byref(vb), self.pod['scales'], _vp(self.pod_vals[l:]), _vp(scr), _st(dev)))
            torch.sum(self.pod_vals.double() * self.pod_coef, 0, out=self._pod64())
            self.terms[self.pod['slot']].copy_(self._pod64())
        for ev in joins:
            cur.wait_event(ev)
        # ---- backward (+ bucketed all-reduce) ---------------------------------------------------------------------------
        if self.ddp is not None:
            arr = self._bucket_struct()
            _lib.check(lib.b2_unet_backward_buckets(plan.handle, self.param_ptrs, self.dl_ptrs, ws, self.grad_ptrs, self.has,
                                                    arr, len(self.buckets), _st(dev)))
            for b in self.buckets:
                self.comm.wait_event(b['em'])
                self.comm.wait_event(b['es'])
                with torch.cuda.stream(self.comm):
                    self.ddp.allreduce_mean_(b['view'])
                    b['done'].record(self.comm)
            for b in self.buckets:
                cur.wait_event(b['done'])
        else:
            _lib.check(lib.b2_unet_backward(plan.handle, self.param_ptrs, self.dl_ptrs, ws, self.grad_ptrs, self.has, _st(dev)))
        # ---- penalty (identical on every rank: added after the all-reduce) ---------------------------------------------------
        for pe in self.pen:
            t = pe['table']
            _lib.check(lib.b2_quadpen_dev(_vp(t.blob), t.n, t.nblocks, pe['coef'], _vp(self.terms[pe['slot']:]), _vp(t.part), _st(dev)))
        if self.sgd is None:
            self._build_sgd()
        t = self.sgd
        _lib.check(lib.b2_sgd_clip_step_dev(_vp(t.blob), t.n, t.nblocks, _vp(self.hyper), self.nesterov, _vp(self.norm), _vp(t.part), _st(dev)))
        torch.sum(self.terms, 0, out=self.total)

    def _pod64(self):
        if not hasattr(self, "_pod_total"):
            self._pod_total = torch.zeros((), dtype=torch.float64, device=self.dev)
        return self._pod_total

    def _level_loss(self, i, x, base):
        lib, dev = self.lib, self.dev
        B, Cc = int(x.shape[0]), int(x.shape[1])
        V = x[0, 0].numel()
        w = self.weights[i]
        slot = self.terms[self.slot_level[i]:]
        scr = self.loss_scr[i]
        y = self.targets[i]
        if base == 'plop':
            thr, max_ent = self.spec['plop']
            D, H, W = (int(s) for s in x.shape[2:])
            if B < 2:
                raise IndexError("too many indices for tensor of dimension 3 (reference PLOP needs B >= 2 on 3D data, Q10)")
            _lib.check(lib.b2_plop_pseudo(_vp(x), _vp(self.tlogits[i]), _vp(y), B, Cc, D, H, W, _vp(thr[i]), float(max_ent), w,
                                          _vp(self.dl[i]), _vp(slot), _vp(scr), _st(dev)))
            return
        cfg = self.spec['cfg']
        ce_only = base == 'ce255'
        _lib.check(lib.b2_dsloss_fwd_bwd(_vp(x), _vp(y), B, Cc, V, w, 0 if ce_only else int(cfg['batch_dice']),
                                         0. if ce_only else float(cfg['smooth']), 0 if ce_only else int(cfg['do_bg']),
                                         255 if ce_only else int(cfg['ignore_index']), 0 if ce_only else 1,
                                         _vp(self.dl[i]), _vp(slot), _vp(scr), _st(dev)))
        if self.spec.get('mib') and w != 0:
            alpha, lkd = self.spec['mib']
            # the KD finalize accumulates into its own slot from several levels / streams: give every level its slot share
            _lib.check(lib.b2_kd_mib(_vp(x), _vp(self.tlogits[i]), B, Cc, V, float(alpha), float(w * lkd), _vp(self.dl[i]),
                                     _vp(slot), _vp(scr), _st(dev)))

    # ------------------------------------------------------------------------------------------------------------
    def _ensure_grad_views(self):
        """param.grad = persistent arena views (an intermixed autograd iteration / zero_grad may have replaced them)"""
        if self.sgd is None:
            return
        k0 = self.sgd_sel[0]
        g = self.params[k0].grad
        if g is None or g.data_ptr() != self.grads[k0].data_ptr():
            sel = set(self.sgd_sel)
            for k, p in enumerate(self.params):
                p.grad = self.grads[k] if k in sel else None

    def run(self, use_graph, targets_ready=None):
        """enqueue (or replay) one iteration on the current stream.  The input patch must already be in self.data (stream
        order); `targets_ready` (event, optional) guards self.targets, which are first read after the forward pass."""
        self.set_hyper()
        self._ensure_grad_views()
        cur = torch.cuda.current_stream(self.dev)
        capture = use_graph and self.graph is None and self.steps_run >= 2   # two eager warm-up iterations first
        if use_graph and self.graph is not None:
            self.graph[0].replay()
            if targets_ready is not None:
                cur.wait_event(targets_ready)
            self.graph[1].replay()
        elif capture:
            if targets_ready is not None:
                cur.wait_event(targets_ready)
            ga, gb = torch.cuda.CUDAGraph(), torch.cuda.CUDAGraph()
            # thread-local capture mode: other host threads (e.g. the GPU patch pipeline's producer thread, which allocates and
            # launches on its own stream) may keep issuing CUDA calls while this thread captures
            with torch.cuda.graph(ga, capture_error_mode="thread_local"):
                self.enqueue_forward()
            with torch.cuda.graph(gb, pool=ga.pool(), capture_error_mode="thread_local"):
                self.enqueue_rest()
            self.graph = (ga, gb)
            ga.replay()
            gb.replay()
        else:
            self.enqueue_forward()
            if targets_ready is not None:
                cur.wait_event(targets_ready)
            self.enqueue_rest()
        self.steps_run += 1
        return self.total


def _ptr_array(tensors):
    arr = (C.c_void_p * len(tensors))()
    for i, t in enumerate(tensors):
        arr[i] = None if t is None else t.data_ptr()
    return arr
