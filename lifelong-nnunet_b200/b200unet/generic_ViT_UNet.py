"""``Generic_ViT_UNet`` -- drop-in for ``nnunet_ext.network_architecture.generic_ViT_UNet.Generic_ViT_UNet`` (reference
generic_ViT_UNet.py:16-338), versions V1, V2 and V3: the convolutional encoder / decoder run in the hand-written sm_100a CUDA
plan (partial passes of include/b2unet.h: b2_unet_forward_parts / b2_unet_backward_parts); the ViT between them runs in the
hand-written kernels of csrc/vit.cu in bf16 mode and through ATen in the fp32 parity mode (b200unet/vision_transformer.py).
V2 / V3 (:299-338) build the ViT input from the first skip plus the bottleneck (V3: plus every skip) up-sampled through the
chain of `tu` transposed convolutions, applied here as standalone ops (b2_tconv3d_fwd / b2_tconv3d_bwd); in these versions
the bottleneck convolutions are live and receive gradients.  V4 (a ViT after every decoder level, with a head as large as the
feature map) is not built.

Data flow of one step (no host copies, no torch.cat):
  encoder part  -> first skip = strided channels-last view of the plan's concat buffer  -> ViT  -> written over the
  bottleneck activation of the plan -> decoder part -> logits;   backward: decoder part leaves d(bottleneck) in the
  plan's gradient arena -> ViT backward (autograd) -> added into the first skip's gradient slice -> encoder part.
V1 discards the bottleneck convolutions' result (generic_ViT_UNet.py:230-253): they are not executed unless a forward
hook sits on them, and their parameters get ``grad = None`` exactly as under autograd in the reference.
"""
import ctypes as C
import math

import torch
from torch import nn

from . import _lib
from .generic_UNet import Generic_UNet, InitWeights_He, softmax_helper, _ptr_array, _NullCtx  # noqa: F401
from .vision_transformer import VisionTransformer, VIT_TYPES

PART_ENCODER, PART_BOTTLENECK, PART_DECODER = 1, 2, 4


def commDiv(a, b):
    """helpful_functions.py:272-286."""
    n = math.gcd(a, b)
    return [i for i in range(1, n + 1) if n % i == 0]


def _view(plan, block, which):
    """NCDHW-shaped strided torch view of an activation (which=1) / gradient (which=2) buffer of conv block `block`."""
    v = _lib.ActView()
    _lib.check(plan.lib.b2_unet_debug_view(plan.handle, C.c_void_p(plan.workspace.data_ptr()), block, which, C.byref(v)))
    esz = 4 if v.dtype == _lib.B2_F32 else 2
    tdt = torch.float32 if v.dtype == _lib.B2_F32 else torch.bfloat16
    off = (v.ptr - plan.workspace.data_ptr()) // esz
    return plan.workspace.view(tdt).as_strided(
        (v.n, v.c, v.d, v.h, v.w), (v.d * v.h * v.w * v.pitch, 1, v.h * v.w * v.pitch, v.w * v.pitch, v.pitch), off)


class _EncoderFunction(torch.autograd.Function):
    """Encoder stages of the plan.  Returns (first skip view, ordering token).  Its backward runs LAST (autograd waits
    for the gradients of both outputs) and hands out the gradients of every U-Net parameter."""

    @staticmethod
    def forward(ctx, net, plan, x, with_bottleneck, *params):
        stream = torch.cuda.current_stream(x.device).cuda_stream
        plan.generation += 1
        parts = PART_ENCODER | (PART_BOTTLENECK if with_bottleneck else 0)
        _lib.check(plan.lib.b2_unet_forward_parts(plan.handle, _ptr_array(params), C.c_void_p(x.data_ptr()),
                                                  C.c_void_p(plan.workspace.data_ptr()), None, parts, C.c_void_p(stream)))
        ctx.plan, ctx.generation, ctx.params = plan, plan.generation, params
        ctx.set_materialize_grads(False)
        skip0 = _view(plan, 1, 1)
        token = torch.zeros((), dtype=torch.float32, device=x.device)
        return skip0, token

    @staticmethod
    def backward(ctx, dskip0, dtoken):
        plan, params = ctx.plan, ctx.params
        if plan.generation != ctx.generation:
            raise RuntimeError("b200unet: workspace overwritten by a later forward before backward ran")
        st = getattr(plan, "_bwd_state", None)
        if st is None:
            raise RuntimeError("b200unet: encoder backward reached before the decoder backward")
        plan._bwd_state = None
        flat, grads, has = st
        dev = params[0].device
        stream = torch.cuda.current_stream(dev).cuda_stream
        if dskip0 is not None:
            _view(plan, 1, 2).add_(dskip0)
        _lib.check(plan.lib.b2_unet_backward_parts(plan.handle, _ptr_array(params), None,
                                                   C.c_void_p(plan.workspace.data_ptr()), _ptr_array(grads), has,
                                                   PART_ENCODER, C.c_void_p(stream)))
        plan.last_flat_grad = flat
        return (None, None, None, None) + tuple(g if has[i] else None for i, g in enumerate(grads))


class _DecoderFunction(torch.autograd.Function):
    """Writes the ViT result over the plan's bottleneck activation and runs the decoder part."""

    @staticmethod
    def forward(ctx, net, plan, vit_out, token, *params):
        stream = torch.cuda.current_stream(vit_out.device).cuda_stream
        bott = _view(plan, 2 * net.num_pool + 1, 1)
        bott.copy_(vit_out.reshape(bott.shape))            # x.reshape(size), generic_ViT_UNet.py:253
        logits = [torch.empty(s, dtype=torch.float32, device=vit_out.device) for s in plan.out_shapes]
        _lib.check(plan.lib.b2_unet_forward_parts(plan.handle, _ptr_array(params), None,
                                                  C.c_void_p(plan.workspace.data_ptr()), _ptr_array(logits),
                                                  PART_DECODER, C.c_void_p(stream)))
        ctx.plan, ctx.generation, ctx.params = plan, plan.generation, params
        ctx.vit_shape, ctx.vit_dtype, ctx.block = vit_out.shape, vit_out.dtype, 2 * net.num_pool + 1
        ctx.set_materialize_grads(False)
        return tuple(logits)

    @staticmethod
    def backward(ctx, *dlogits):
        plan, params = ctx.plan, ctx.params
        if plan.generation != ctx.generation:
            raise RuntimeError("b200unet: workspace overwritten by a later forward before backward ran")
        dev = params[0].device
        stream = torch.cuda.current_stream(dev).cuda_stream
        dl = [None if d is None else d.contiguous().float() for d in dlogits]
        flat = torch.empty(sum(plan.param_numel), dtype=torch.float32, device=dev)
        grads, o = [], 0
        for n_, s in zip(plan.param_numel, plan.param_shapes):
            grads.append(flat[o:o + n_].view(s))
            o += n_
        has = (C.c_int32 * len(grads))()
        _lib.check(plan.lib.b2_unet_backward_parts(plan.handle, _ptr_array(params), _ptr_array(dl),
                                                   C.c_void_p(plan.workspace.data_ptr()), _ptr_array(grads), has,
                                                   PART_DECODER, C.c_void_p(stream)))
        plan._bwd_state = (flat, grads, has)
        dvit = _view(plan, ctx.block, 2).to(ctx.vit_dtype).reshape(ctx.vit_shape)
        return (None, None, dvit, torch.zeros((), dtype=torch.float32, device=dev)) + (None,) * len(params)


def _cl_strides_ok(t):
    st, (B, Cc, D, H, W) = t.stride(), t.shape
    return st[1] == 1 and st[4] >= Cc and st[3] == W * st[4] and st[2] == H * st[3] and st[0] == D * st[2]


class _TconvFunction(torch.autograd.Function):
    """y = ConvTranspose3d(kernel == stride, bias=False)(x) on channels-last (NDHWC) tensors through b2_tconv3d_fwd / _bwd --
    the `tu` modules applied outside the decoder (generic_ViT_UNet.py:306-308, 323-334).  x may be a strided view of the plan's
    workspace (concat buffers have pitch 2C)."""

    @staticmethod
    def _desc(x, w, k, out_pitch):
        d = _lib.TconvDesc()
        d.n, d.cin, d.d, d.h, d.w = (int(v) for v in x.shape)
        d.cout = int(w.shape[1])
        d.k = (C.c_int32 * 3)(*k)
        d.in_pitch, d.out_pitch = int(x.stride(4)), int(out_pitch)
        d.dtype = _lib.B2_F32 if x.dtype == torch.float32 else _lib.B2_BF16
        return d

    @staticmethod
    def forward(ctx, x, w, k):
        lib = _lib.load()
        if not _cl_strides_ok(x) or x.dtype not in (torch.float32, torch.bfloat16):
            raise RuntimeError("b2_tconv3d needs a channels-last fp32 / bf16 tensor")
        k = tuple(int(v) for v in k)
        assert tuple(w.shape[2:]) == k and w.dtype == torch.float32 and w.is_contiguous()
        B, _, D, H, W = x.shape
        y = torch.empty((B, int(w.shape[1]), D * k[0], H * k[1], W * k[2]), dtype=x.dtype, device=x.device,
                        memory_format=torch.channels_last_3d)
        d = _TconvFunction._desc(x, w, k, w.shape[1])
        scr = torch.empty(int(lib.b2_tconv3d_scratch_bytes(C.byref(d))), dtype=torch.uint8, device=x.device)
        st = C.c_void_p(torch.cuda.current_stream(x.device).cuda_stream)
        _lib.check(lib.b2_tconv3d_fwd(C.byref(d), C.c_void_p(x.data_ptr()), C.c_void_p(w.data_ptr()), C.c_void_p(y.data_ptr()),
                                      C.c_void_p(scr.data_ptr()), st))
        # (not save_for_backward: x may be one of several views of the plan's workspace, whose shared version counter moves when
        # the backward pass accumulates into other views; the plan's generation counter guards the data instead)
        ctx.x, ctx.w, ctx.k = x.detach(), w, k
        return y

    @staticmethod
    def backward(ctx, dy):
        lib = _lib.load()
        x, w = ctx.x, ctx.w
        dy = dy.to(x.dtype).contiguous(memory_format=torch.channels_last_3d)
        d = _TconvFunction._desc(x, w, ctx.k, dy.stride(4))
        need_dx, need_dw = ctx.needs_input_grad[0], ctx.needs_input_grad[1]
        dx = torch.empty(tuple(x.shape), dtype=x.dtype, device=x.device, memory_format=torch.channels_last_3d) if need_dx else None
        dw = torch.empty_like(w) if need_dw else None
        d.in_pitch = int(x.stride(4))
        dxd = _lib.TconvDesc.from_buffer_copy(d)
        scr = torch.empty(int(lib.b2_tconv3d_scratch_bytes(C.byref(d))), dtype=torch.uint8, device=x.device)
        st = C.c_void_p(torch.cuda.current_stream(x.device).cuda_stream)
        if need_dw:      # weight gradient reads x at its own pitch
            _lib.check(lib.b2_tconv3d_bwd(C.byref(d), C.c_void_p(x.data_ptr()), C.c_void_p(dy.data_ptr()), C.c_void_p(w.data_ptr()),
                                          None, C.c_void_p(dw.data_ptr()), C.c_void_p(scr.data_ptr()), st))
        if need_dx:      # data gradient is written dense (pitch = Cin)
            dxd.in_pitch = int(x.shape[1])
            _lib.check(lib.b2_tconv3d_bwd(C.byref(dxd), C.c_void_p(x.data_ptr()), C.c_void_p(dy.data_ptr()), C.c_void_p(w.data_ptr()),
                                          C.c_void_p(dx.data_ptr()), None, C.c_void_p(scr.data_ptr()), st))
        return dx, dw, None


class _EncoderAllFunction(torch.autograd.Function):
    """Encoder stages AND the bottleneck of the plan (V2 / V3).  Returns (skip_0 .. skip_{P-1} as views of the plan's workspace,
    a copy of the bottleneck output -- the decoder part overwrites that buffer with the ViT result --, ordering token).  Its
    backward runs last: the gradients autograd hands back for the skips are added into the plan's gradient buffers (which already
    hold the decoder's contributions), the bottleneck's replaces what the decoder left there (that was d(ViT output))."""

    @staticmethod
    def forward(ctx, net, plan, x, *params):
        stream = torch.cuda.current_stream(x.device).cuda_stream
        plan.generation += 1
        _lib.check(plan.lib.b2_unet_forward_parts(plan.handle, _ptr_array(params), C.c_void_p(x.data_ptr()),
                                                  C.c_void_p(plan.workspace.data_ptr()), None, PART_ENCODER | PART_BOTTLENECK,
                                                  C.c_void_p(stream)))
        ctx.plan, ctx.generation, ctx.params, ctx.P = plan, plan.generation, params, net.num_pool
        ctx.set_materialize_grads(False)
        skips = [_view(plan, 2 * d + 1, 1) for d in range(net.num_pool)]
        bott = _view(plan, 2 * net.num_pool + 1, 1).clone()
        token = torch.zeros((), dtype=torch.float32, device=x.device)
        return tuple(skips) + (bott, token)

    @staticmethod
    def backward(ctx, *douts):
        plan, params, P = ctx.plan, ctx.params, ctx.P
        if plan.generation != ctx.generation:
            raise RuntimeError("b200unet: workspace overwritten by a later forward before backward ran")
        st = getattr(plan, "_bwd_state", None)
        if st is None:
            raise RuntimeError("b200unet: encoder backward reached before the decoder backward")
        plan._bwd_state = None
        flat, grads, has = st
        dev = params[0].device
        stream = torch.cuda.current_stream(dev).cuda_stream
        for d in range(P):
            if douts[d] is not None:
                _view(plan, 2 * d + 1, 2).add_(douts[d])
        dbott = _view(plan, 2 * P + 1, 2)
        if douts[P] is not None:
            dbott.copy_(douts[P])
        else:
            dbott.zero_()
        _lib.check(plan.lib.b2_unet_backward_parts(plan.handle, _ptr_array(params), None,
                                                   C.c_void_p(plan.workspace.data_ptr()), _ptr_array(grads), has,
                                                   PART_ENCODER | PART_BOTTLENECK, C.c_void_p(stream)))
        plan.last_flat_grad = flat
        return (None, None, None) + tuple(g if has[i] else None for i, g in enumerate(grads))


class Generic_ViT_UNet(Generic_UNet):
    def __init__(self, input_channels, base_num_features, num_classes, num_pool, patch_size, num_conv_per_stage=2,
                 feat_map_mul_on_downscale=2, conv_op=nn.Conv3d, norm_op=nn.InstanceNorm3d, norm_op_kwargs=None,
                 dropout_op=nn.Dropout3d, dropout_op_kwargs=None, nonlin=nn.LeakyReLU, nonlin_kwargs=None,
                 deep_supervision=True, dropout_in_localization=False, final_nonlin=lambda x: x,
                 weightInitializer=InitWeights_He(1e-2), pool_op_kernel_sizes=None, conv_kernel_sizes=None,
                 upscale_logits=False, convolutional_pooling=True, convolutional_upsampling=True,
                 max_num_features=None, basic_block=None, seg_output_use_bias=False,
                 vit_version='V1', vit_type='base', split_gpu=False, ViT_task_specific_ln=False, first_task_name=None,
                 do_LSA=False, do_SPT=False):
        super().__init__(input_channels, base_num_features, num_classes, num_pool, num_conv_per_stage,
                         feat_map_mul_on_downscale, conv_op, norm_op, norm_op_kwargs, dropout_op, dropout_op_kwargs,
                         nonlin, nonlin_kwargs, deep_supervision, dropout_in_localization, final_nonlin,
                         weightInitializer, pool_op_kernel_sizes, conv_kernel_sizes, upscale_logits,
                         convolutional_pooling, convolutional_upsampling, max_num_features, basic_block,
                         seg_output_use_bias)
        assert isinstance(patch_size, list) and all(isinstance(n, int) for n in patch_size), \
            'Please provide the patch_size in form of a list of integers..'
        if len(patch_size) != 3:
            raise NotImplementedError("b200unet.Generic_ViT_UNet is 3D only")
        vit_type = vit_type.lower()
        assert vit_type in VIT_TYPES, "Please provide one of the following three types: 'base', 'large' or 'huge'. " \
                                      "You provided '{}'".format(vit_type)
        self.version = vit_version.title()
        assert self.version in ['V1', 'V2', 'V3', 'V4'], 'Please provide a correct version (V1, V2, V3 or V4), not {}.'.format(vit_version)
        if self.version == 'V4' or split_gpu or do_SPT:
            raise NotImplementedError("b200unet.Generic_ViT_UNet: vit_version V1 / V2 / V3 on one device (optionally with LSA or "
                                      "task-specific LayerNorms) are implemented; V4, SPT and split_gpu are not (no eager fallback)")
        self.prepare = {'V1': '_get_ViT_inputV1', 'V2': '_get_ViT_inputV2', 'V3': '_get_ViT_inputV3'}
        self.split_gpu, self.use_skip = False, 0
        self.ViT_types = VIT_TYPES
        # sizes the reference obtains from a dry run (generic_ViT_UNet.py:85-131) follow from the pooling geometry
        dhw = [int(s) for s in patch_size]
        self.skip_sizes = []
        feats = [min(base_num_features * 2 ** d, self.max_num_features) for d in range(num_pool + 1)]
        for d in range(num_pool + 1):
            if d > 0:
                dhw = [(s - 1) // k + 1 for s, k in zip(dhw, self.pool_op_kernel_sizes[d - 1])]
            if d < num_pool:
                self.skip_sizes.append(torch.Size([1, feats[d]] + dhw))
        self.num_classesViT = int(feats[num_pool] * dhw[0] * dhw[1] * dhw[2])
        self.img_size = list(self.skip_sizes[0][2:])
        patch_dim = max(x for x in commDiv(self.img_size[0], self.img_size[1]) if x <= 16)   # :148
        self.patch_size = (patch_dim, patch_dim)
        self.in_chans = int(self.skip_sizes[0][1])
        cfg = VIT_TYPES[vit_type]
        vit = VisionTransformer(ViT_2d=False, img_size=self.img_size, patch_size=self.patch_size,
                                img_depth=[self.img_size[0]], in_chans=self.in_chans, num_classes=self.num_classesViT,
                                embed_dim=cfg['embed_size'], depth=cfg['layers'], num_heads=cfg['head'], mlp_ratio=4,
                                qkv_bias=True, task_specific_ln=ViT_task_specific_ln, task_name=first_task_name, is_LSA=do_LSA)
        # registration order of generic_ViT_UNet.py:193-211
        parts = {n: getattr(self, n) for n in ('conv_blocks_localization', 'conv_blocks_context', 'td', 'tu', 'seg_outputs')}
        for n in parts:
            delattr(self, n)
        for n in ('conv_blocks_localization', 'conv_blocks_context', 'ViT', 'td', 'tu', 'seg_outputs'):
            setattr(self, n, vit if n == 'ViT' else parts[n])
        self.split_names = ['ViT']

    def forward(self, x, store_vit_input=False):
        if not x.is_cuda:
            raise RuntimeError("b200unet.Generic_ViT_UNet runs on a CUDA device only (sm_100a); there is no CPU fallback")
        plan = self._get_plan(x)
        x = x.contiguous().float()
        params = self._ordered_params(plan)
        mods = self._conv_modules(plan)
        P = self.num_pool
        with_bott = any(bool(mods[i]._forward_hooks) for i in (2 * P, 2 * P + 1))
        grad = torch.is_grad_enabled() and (any(p.requires_grad for p in params) or
                                            any(p.requires_grad for p in self.ViT.parameters()))
        dvit_in = None
        if self.version == 'V1':
            if grad:
                skip0, token = _EncoderFunction.apply(self, plan, x, with_bott, *params)
            else:
                with torch.no_grad():
                    skip0, token = _EncoderFunction.forward(_NullCtx(), self, plan, x, with_bott, *params)
            if grad:
                dvit_in = _view(plan, 1, 2)      # the ViT's input gradient is added straight into the first skip's gradient buffer
        else:
            if grad:
                outs_e = _EncoderAllFunction.apply(self, plan, x, *params)
            else:
                with torch.no_grad():
                    outs_e = _EncoderAllFunction.forward(_NullCtx(), self, plan, x, *params)
            skips, last_context, token = list(outs_e[:P]), outs_e[P], outs_e[P + 1]
            skip0 = getattr(self, self.prepare[self.version])(skips, last_context).contiguous(memory_format=torch.channels_last_3d)
        if store_vit_input:
            self.ViT_in = skip0.clone()
        if self.precision != "fp32" and self.ViT.native_supported(skip0):
            # bf16 mode: the whole ViT runs in the hand-written kernels of csrc/vit.cu; its input gradient is added straight
            # into the plan's gradient buffer of the first skip (so the encoder backward needs no extra add)
            bott = _view(plan, 2 * P + 1, 1)
            vit_out = self.ViT.forward_native(skip0, dvit_in, tuple(int(v) for v in bott.shape[1:]),
                                              return_input_grad=grad and self.version != 'V1')
        else:
            if self.precision != "fp32" and not self.ViT.store_attn_weights:
                # no silent library fallback: the bf16 production mode either runs the hand-written ViT or says why it cannot
                raise NotImplementedError("native ViT kernels (csrc/vit.cu) need head dim 64, embed <= 1024, batch <= 4 and a "
                                          "channel count that is a multiple of 8; use precision='fp32' (ATen parity mode) for "
                                          "other shapes -- got batch %d, %d channels, embed %d" %
                                          (skip0.shape[0], skip0.shape[1], self.ViT.embed_dim))
            # fp32 parity mode (and `store_attn_weights`, which asks for the materialised probabilities): the same module
            # evaluated through ATen
            with torch.autocast('cuda', dtype=torch.bfloat16, enabled=self.precision != "fp32"):
                vit_out = self.ViT(skip0)
        if grad:
            outs = _DecoderFunction.apply(self, plan, vit_out, token, *params)
        else:
            with torch.no_grad():
                outs = _DecoderFunction.forward(_NullCtx(), self, plan, vit_out, token, *params)
        self._last_plan = plan
        self._fire_hooks(plan, outs)
        outs = tuple(self.final_nonlin(o) for o in outs)
        if self._deep_supervision and self.do_ds:
            return outs
        return outs[0]

    # -- ViT input preparation (generic_ViT_UNet.py:290-338); V1 is handled inline in forward --------------------------------
    def _tu(self, u, t):
        k = self.pool_op_kernel_sizes[-(u + 1)]
        return _TconvFunction.apply(t, self.tu[u].weight, tuple(int(v) for v in k))

    def _get_ViT_inputV2(self, skips, last_context):
        """first skip + the bottleneck output up-sampled through every `tu` (no skip concatenation, no convolutions)"""
        t = last_context
        for u in range(len(self.tu)):
            t = self._tu(u, t)
        return skips[self.use_skip] + t

    def _get_ViT_inputV3(self, skips, last_context):
        """the up-sampled bottleneck plus EVERY skip up-sampled to full resolution (the first skip as it is)"""
        t = last_context
        for u in range(len(self.tu)):
            t = self._tu(u, t)
        vit_in = t
        for idx, skip in enumerate(reversed(skips)):
            t = skip
            for u in range(idx + 1, len(self.tu)):
                t = self._tu(u, t)
            vit_in = vit_in + t
        return vit_in
