"""``Generic_UNet`` -- drop-in for ``nnunet.network_architecture.generic_UNet.Generic_UNet`` whose arithmetic runs in
hand-written sm_100a CUDA behind the C ABI (include/b2unet.h).

Plugin surface kept:
* constructor signature and positional order of the upstream class (reference
  nnunet_ext/network_architecture/generic_UNet.py:17-35; call site restated at
  nnunet_ext/training/network_training/nnViTUNetTrainer.py:101-125);
* module tree / ``state_dict`` keys ``conv_blocks_context[i].blocks[j].{conv,instnorm,lrelu}``, ``td``, ``tu[i]``,
  ``conv_blocks_localization[i][0|1].blocks[0].{...}``, ``seg_outputs[i]`` (fixture: reference
  test/network_architecture/test_MultiHead_Module.py:281-432) with plain fp32 ``nn.Parameter`` ownership, so
  ``MultiHead_Module`` splitting / deepcopy / ``load_state_dict`` (reference MultiHead_Module.py:139-157,326-377),
  checkpoints and ``named_parameters()``-keyed Fisher dicts keep working;
* ``forward(x)`` returns the deep-supervision tuple (highest resolution first) when ``do_ds`` else the full-res tensor
  (reference generic_ViT_UNet.py:280-286);
* forward hooks registered on the conv sub-modules (PLOP: reference plop:330-353) are fired with the raw conv outputs.

Only the configuration the trainers build is supported (3D, InstanceNorm3d, LeakyReLU, dropout p=0, conv pooling and
conv up-sampling, 3x3x3 kernels, 2 convs per stage, identity final nonlinearity); anything else raises -- there is no
eager / CPU fallback.
"""
import ctypes as C

import numpy as np
import torch
from torch import nn

from . import _lib


class InitWeights_He(object):
    """kaiming_normal_(a=neg_slope) on conv weights, zero bias."""

    def __init__(self, neg_slope=1e-2):
        self.neg_slope = neg_slope

    def __call__(self, module):
        if isinstance(module, (nn.Conv3d, nn.Conv2d, nn.ConvTranspose2d, nn.ConvTranspose3d)):
            module.weight = nn.init.kaiming_normal_(module.weight, a=self.neg_slope)
            if module.bias is not None:
                module.bias = nn.init.constant_(module.bias, 0)


def softmax_helper(x):
    return torch.softmax(x, 1)


class ConvDropoutNormNonlin(nn.Module):
    """Parameter container with the upstream attribute names (conv / dropout / instnorm / lrelu)."""

    def __init__(self, cin, cout, stride, norm_kwargs, nonlin_kwargs):
        super().__init__()
        self.conv = nn.Conv3d(cin, cout, kernel_size=[3, 3, 3], stride=tuple(stride), padding=[1, 1, 1], bias=True)
        self.dropout = None
        self.instnorm = nn.InstanceNorm3d(cout, **norm_kwargs)
        self.lrelu = nn.LeakyReLU(**nonlin_kwargs)

    def forward(self, x):  # pragma: no cover - the fused path never calls sub-modules
        raise RuntimeError("b200unet: sub-modules are parameter containers; call the Generic_UNet itself")


class StackedConvLayers(nn.Module):
    def __init__(self, cin, cout, num_convs, first_stride, norm_kwargs, nonlin_kwargs):
        super().__init__()
        self.input_channels, self.output_channels = cin, cout
        blocks = [ConvDropoutNormNonlin(cin, cout, first_stride or (1, 1, 1), norm_kwargs, nonlin_kwargs)]
        blocks += [ConvDropoutNormNonlin(cout, cout, (1, 1, 1), norm_kwargs, nonlin_kwargs) for _ in range(num_convs - 1)]
        self.blocks = nn.Sequential(*blocks)

    def forward(self, x):  # pragma: no cover
        raise RuntimeError("b200unet: sub-modules are parameter containers; call the Generic_UNet itself")


class _Plan:
    """One b2_unet_plan + its HBM workspace, for a fixed (batch, patch, dtype)."""

    def __init__(self, net, batch, patch, device, act_dtype):
        lib = _lib.load()
        g = _lib.Geometry()
        g.batch, g.in_channels, g.num_classes = batch, net.input_channels, net.num_classes
        g.base_features, g.max_features, g.num_pool = net.base_num_features, net.max_num_features, net.num_pool
        for i in range(3):
            g.patch[i] = int(patch[i])
        for l, k in enumerate(net.pool_op_kernel_sizes):
            for i in range(3):
                g.pool[l][i] = int(k[i])
        g.act_dtype = act_dtype
        g.lrelu_slope = float(net.nonlin_kwargs['negative_slope'])
        g.norm_eps = float(net.norm_op_kwargs['eps'])
        h = C.c_void_p()
        _lib.check(lib.b2_unet_plan_create(C.byref(g), C.byref(h)))
        self.lib, self.handle, self.device, self.act_dtype = lib, h, device, act_dtype
        self.batch, self.patch = batch, tuple(int(p) for p in patch)
        n = lib.b2_unet_num_params(h)
        self.param_names, self.param_shapes = [], []
        info = _lib.ParamInfo()
        for i in range(n):
            _lib.check(lib.b2_unet_param_info(h, i, C.byref(info)))
            self.param_names.append(info.name.decode())
            self.param_shapes.append(tuple(int(info.shape[j]) for j in range(info.ndim)))
        self.param_numel = [int(np.prod(s)) for s in self.param_shapes]
        self.out_shapes = []
        dhw = (C.c_int32 * 3)()
        for lvl in range(net.num_pool):
            _lib.check(lib.b2_unet_output_shape(h, lvl, C.byref(dhw)))
            self.out_shapes.append((batch, net.num_classes, dhw[0], dhw[1], dhw[2]))
        self.workspace_bytes = int(lib.b2_unet_workspace_bytes(h))
        self.workspace = torch.empty(self.workspace_bytes, dtype=torch.uint8, device=device)
        self.generation = 0
        self.conv_names = []
        buf = C.create_string_buffer(96)
        for i in range(lib.b2_unet_num_convs(h)):
            _lib.check(lib.b2_unet_conv_name(h, i, buf))
            self.conv_names.append(buf.value.decode())

    def __del__(self):
        try:
            self.lib.b2_unet_plan_destroy(self.handle)
        except Exception:
            pass

    def conv_output(self, idx):
        """NCDHW-shaped (channels-last-3d strided) torch view of raw conv output `idx` inside the workspace."""
        v = _lib.ActView()
        _lib.check(self.lib.b2_unet_conv_output(self.handle, C.c_void_p(self.workspace.data_ptr()), idx, C.byref(v)))
        esz = 4 if v.dtype == _lib.B2_F32 else 2
        tdt = torch.float32 if v.dtype == _lib.B2_F32 else torch.bfloat16
        off = (v.ptr - self.workspace.data_ptr()) // esz
        flat = self.workspace.view(tdt)
        return flat.as_strided((v.n, v.c, v.d, v.h, v.w),
                               (v.d * v.h * v.w * v.pitch, 1, v.h * v.w * v.pitch, v.w * v.pitch, v.pitch), off), v


def _ptr_array(tensors):
    arr = (C.c_void_p * len(tensors))()
    for i, t in enumerate(tensors):
        arr[i] = None if t is None else t.data_ptr()
    return arr


class _UNetFunction(torch.autograd.Function):
    @staticmethod
    def forward(ctx, net, plan, x, *params):
        stream = torch.cuda.current_stream(x.device).cuda_stream
        logits = [torch.empty(s, dtype=torch.float32, device=x.device) for s in plan.out_shapes]
        plan.generation += 1
        _lib.check(plan.lib.b2_unet_forward(plan.handle, _ptr_array(params), C.c_void_p(x.data_ptr()),
                                            C.c_void_p(plan.workspace.data_ptr()), _ptr_array(logits), 1,
                                            C.c_void_p(stream)))
        ctx.plan, ctx.generation, ctx.params = plan, plan.generation, params
        ctx.set_materialize_grads(False)
        return tuple(logits)

    @staticmethod
    def backward(ctx, *dlogits):
        plan, params = ctx.plan, ctx.params
        if plan.generation != ctx.generation:
            raise RuntimeError("b200unet: the workspace was overwritten by a later forward of the same network "
                               "before backward ran; run forward/backward pairs back to back")
        dev = params[0].device
        stream = torch.cuda.current_stream(dev).cuda_stream
        dl = [None if d is None else d.contiguous().float() for d in dlogits]
        total = sum(plan.param_numel)
        flat = torch.empty(total, dtype=torch.float32, device=dev)
        grads, o = [], 0
        for n_, s in zip(plan.param_numel, plan.param_shapes):
            grads.append(flat[o:o + n_].view(s))
            o += n_
        has = (C.c_int32 * len(grads))()
        _lib.check(plan.lib.b2_unet_backward(plan.handle, _ptr_array(params), _ptr_array(dl),
                                             C.c_void_p(plan.workspace.data_ptr()), _ptr_array(grads), has,
                                             C.c_void_p(stream)))
        plan.last_flat_grad = flat
        out = tuple(g if has[i] else None for i, g in enumerate(grads))
        return (None, None, None) + out


class Generic_UNet(nn.Module):
    DEFAULT_BATCH_SIZE_3D = 2
    DEFAULT_PATCH_SIZE_3D = (64, 192, 160)
    SPACING_FACTOR_BETWEEN_STAGES = 2
    BASE_NUM_FEATURES_3D = 30
    MAX_NUMPOOL_3D = 999
    MAX_NUM_FILTERS_3D = 320

    def __init__(self, input_channels, base_num_features, num_classes, num_pool, num_conv_per_stage=2,
                 feat_map_mul_on_downscale=2, conv_op=nn.Conv3d,
                 norm_op=nn.InstanceNorm3d, norm_op_kwargs=None,
                 dropout_op=nn.Dropout3d, dropout_op_kwargs=None,
                 nonlin=nn.LeakyReLU, nonlin_kwargs=None, deep_supervision=True, dropout_in_localization=False,
                 final_nonlin=lambda x: x, weightInitializer=InitWeights_He(1e-2), pool_op_kernel_sizes=None,
                 conv_kernel_sizes=None, upscale_logits=False, convolutional_pooling=True,
                 convolutional_upsampling=True, max_num_features=None, basic_block=None,
                 seg_output_use_bias=False):
        super().__init__()
        if norm_op_kwargs is None:
            norm_op_kwargs = {'eps': 1e-5, 'affine': True}
        if nonlin_kwargs is None:
            nonlin_kwargs = {'negative_slope': 1e-2, 'inplace': True}
        if dropout_op_kwargs is None:
            dropout_op_kwargs = {'p': 0, 'inplace': True}
        if pool_op_kernel_sizes is None:
            pool_op_kernel_sizes = [(2, 2, 2)] * num_pool
        if conv_kernel_sizes is None:
            conv_kernel_sizes = [(3, 3, 3)] * (num_pool + 1)

        def unsupported(what):
            raise NotImplementedError("b200unet.Generic_UNet supports only the configuration built by "
                                      "nnUNetTrainerV2.initialize_network (3D conv, InstanceNorm3d affine, LeakyReLU, "
                                      "dropout p=0, conv pooling + conv upsampling, 3x3x3 kernels, 2 convs/stage); "
                                      "got " + what)
        if conv_op is not nn.Conv3d: unsupported("conv_op=%s" % conv_op)
        if norm_op is not nn.InstanceNorm3d or not norm_op_kwargs.get('affine', False): unsupported("norm_op=%s %s" % (norm_op, norm_op_kwargs))
        if nonlin is not nn.LeakyReLU: unsupported("nonlin=%s" % nonlin)
        if dropout_op is not None and dropout_op_kwargs.get('p', 0) not in (0, None): unsupported("dropout p>0")
        if not (convolutional_pooling and convolutional_upsampling): unsupported("max-pool / interpolation path")
        if num_conv_per_stage != 2 or feat_map_mul_on_downscale != 2: unsupported("num_conv_per_stage/feat_mul")
        if any(tuple(k) != (3, 3, 3) for k in conv_kernel_sizes): unsupported("conv_kernel_sizes=%s" % (conv_kernel_sizes,))
        if upscale_logits or seg_output_use_bias or not deep_supervision: unsupported("upscale_logits/seg bias/no DS")
        if len(pool_op_kernel_sizes) != num_pool or num_pool > 7: unsupported("pool_op_kernel_sizes")
        if any(int(s) not in (1, 2) for k in pool_op_kernel_sizes for s in k): unsupported("pool strides other than 1/2")

        self.input_channels, self.base_num_features, self.num_classes = input_channels, base_num_features, num_classes
        self.num_pool = num_pool
        self.conv_op, self.norm_op, self.norm_op_kwargs = conv_op, norm_op, norm_op_kwargs
        self.dropout_op, self.dropout_op_kwargs = dropout_op, dropout_op_kwargs
        self.nonlin, self.nonlin_kwargs = nonlin, nonlin_kwargs
        self.final_nonlin = final_nonlin
        self.weightInitializer = weightInitializer
        self.convolutional_pooling, self.convolutional_upsampling = True, True
        self.upscale_logits = False
        self._deep_supervision = self.do_ds = deep_supervision
        self.inference_apply_nonlin = lambda x: x
        self.pool_op_kernel_sizes = [tuple(int(s) for s in k) for k in pool_op_kernel_sizes]
        self.conv_kernel_sizes = [tuple(k) for k in conv_kernel_sizes]
        self.max_num_features = self.MAX_NUM_FILTERS_3D if max_num_features is None else max_num_features
        self.input_shape_must_be_divisible_by = np.prod(self.pool_op_kernel_sizes, 0, dtype=np.int64)
        self.precision = "fp32"    # "fp32" (parity mode) | "bf16" (activation storage + tensor-core math)

        feats, f = [], base_num_features
        for _ in range(num_pool + 1):
            feats.append(min(f, self.max_num_features))
            f = int(np.round(f * 2))
        nk, ak = dict(norm_op_kwargs), dict(nonlin_kwargs)
        ctx, cin = [], input_channels
        for d in range(num_pool):
            ctx.append(StackedConvLayers(cin, feats[d], 2, self.pool_op_kernel_sizes[d - 1] if d > 0 else None, nk, ak))
            cin = feats[d]
        ctx.append(nn.Sequential(StackedConvLayers(cin, feats[num_pool], 1, self.pool_op_kernel_sizes[-1], nk, ak),
                                 StackedConvLayers(feats[num_pool], feats[num_pool], 1, None, nk, ak)))
        loc, tu, seg = [], [], []
        down = feats[num_pool]
        for u in range(num_pool):
            skip = feats[num_pool - 1 - u]
            k = self.pool_op_kernel_sizes[-(u + 1)]
            tu.append(nn.ConvTranspose3d(down, skip, k, k, bias=False))
            loc.append(nn.Sequential(StackedConvLayers(2 * skip, skip, 1, None, nk, ak),
                                     StackedConvLayers(skip, skip, 1, None, nk, ak)))
            seg.append(nn.Conv3d(skip, num_classes, 1, 1, 0, 1, 1, False))
            down = skip
        # registration order of the upstream class (fixture :283,345,417,422,427)
        self.conv_blocks_localization = nn.ModuleList(loc)
        self.conv_blocks_context = nn.ModuleList(ctx)
        self.td = nn.ModuleList([])
        self.tu = nn.ModuleList(tu)
        self.seg_outputs = nn.ModuleList(seg)
        self.upscale_logits_ops = [lambda x: x] * (num_pool - 1)
        if self.weightInitializer is not None:
            self.apply(self.weightInitializer)
        self._plans = {}

    # ----------------------------------------------------------------------------------------------------------
    def _get_plan(self, x):
        if x.dim() != 5 or x.shape[1] != self.input_channels:
            raise ValueError("expected input (B,%d,D,H,W), got %s" % (self.input_channels, tuple(x.shape)))
        act = _lib.B2_F32 if self.precision == "fp32" else _lib.B2_BF16
        key = (int(x.shape[0]), tuple(int(s) for s in x.shape[2:]), x.device, act)
        plan = self._plans.get(key)
        if plan is None:
            plan = _Plan(self, key[0], key[1], x.device, act)
            self._plans[key] = plan
        return plan

    def __getstate__(self):
        # plans own C handles and HBM workspaces: never copied / pickled (copy.deepcopy of the network is how the
        # reference creates teachers, mib:90-97 / plop:184-195)
        d = self.__dict__.copy()
        d['_plans'] = {}
        d.pop('_last_plan', None)
        return d

    def _apply(self, fn, *a, **kw):
        self._plans = {}          # device / dtype moves invalidate plans (workspaces live on the old device)
        return super()._apply(fn, *a, **kw)

    def _ordered_params(self, plan):
        cached = getattr(plan, "_params", None)
        if cached is not None:
            return cached
        table = dict(self.named_parameters())
        out = []
        for name, shape in zip(plan.param_names, plan.param_shapes):
            p = table[name]
            if tuple(p.shape) != shape or p.dtype != torch.float32 or not p.is_contiguous():
                raise RuntimeError("parameter %s: expected contiguous fp32 %s, got %s %s" % (name, shape, p.dtype, tuple(p.shape)))
            out.append(p)
        plan._params = out
        return out

    def _conv_modules(self, plan):
        cached = getattr(plan, "_conv_mods", None)
        if cached is None:
            mods = dict(self.named_modules())
            cached = plan._conv_mods = [mods[n] for n in plan.conv_names]
        return cached

    def forward(self, x):
        if not x.is_cuda:
            raise RuntimeError("b200unet.Generic_UNet runs on a CUDA device only (sm_100a); there is no CPU fallback")
        plan = self._get_plan(x)
        x = x.contiguous().float()
        params = self._ordered_params(plan)
        if torch.is_grad_enabled() and any(p.requires_grad for p in params):
            outs = _UNetFunction.apply(self, plan, x, *params)
        else:
            with torch.no_grad():
                outs = _UNetFunction.forward(_NullCtx(), self, plan, x, *params)
        self._last_plan = plan
        self._fire_hooks(plan, outs)
        outs = tuple(self.final_nonlin(o) for o in outs)
        if self._deep_supervision and self.do_ds:
            return outs
        return outs[0]

    def head_logits(self, weight, level=0):
        """Logits of deep-supervision `level` for another ``seg_outputs[num_pool-1-level]`` weight on the decoder
        activation the LAST forward left in the plan's workspace (LwF old-task heads, reference lwf:315-346: the stored
        heads share the body with the running model)."""
        plan = getattr(self, "_last_plan", None)
        if plan is None:
            raise RuntimeError("head_logits: run a forward first")
        w = weight.detach()
        if w.device != plan.workspace.device or w.dtype != torch.float32 or not w.is_contiguous():
            w = w.to(plan.workspace.device, torch.float32).contiguous()
        if w.numel() != self.num_classes * self.seg_outputs[self.num_pool - 1 - level].weight.shape[1]:
            raise ValueError("head weight has %d elements" % w.numel())
        out = torch.empty(plan.out_shapes[level], dtype=torch.float32, device=w.device)
        _lib.check(plan.lib.b2_unet_head_forward(plan.handle, C.c_void_p(plan.workspace.data_ptr()), int(level),
                                                 C.c_void_p(w.data_ptr()), C.c_void_p(out.data_ptr()),
                                                 C.c_void_p(torch.cuda.current_stream(w.device).cuda_stream)))
        return self.final_nonlin(out)

    def _fire_hooks(self, plan, outs):
        """Fire forward hooks of the conv modules with the raw conv outputs (reference plop:330-353) in execution
        order: encoder convs, then per decoder level tu.u, loc.u.0, loc.u.1, seg_outputs.u."""
        mods = self._conv_modules(plan)
        n_enc = 2 * (self.num_pool + 1)
        for i, m in enumerate(mods):
            if m._forward_hooks:
                out_view, _ = plan.conv_output(i)
                for hook in list(m._forward_hooks.values()):
                    hook(m, (None,), out_view)
            if i >= n_enc and (i - n_enc) % 3 == 2:
                u = (i - n_enc) // 3
                sm = self.seg_outputs[u]
                if sm._forward_hooks:
                    for hook in list(sm._forward_hooks.values()):
                        hook(sm, (None,), outs[self.num_pool - 1 - u])


class _NullCtx:
    def set_materialize_grads(self, v):
        pass
