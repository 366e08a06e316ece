"""``Generic_ViT_UNet`` -- drop-in for ``nnunet_ext.network_architecture.generic_ViT_UNet.Generic_ViT_UNet`` (reference
generic_ViT_UNet.py:16-338), version V1: the convolutional encoder / decoder run in the hand-written sm_100a CUDA plan
(partial passes of include/b2unet.h: b2_unet_forward_parts / b2_unet_backward_parts), the ViT between them runs through
ATen on the same device (b200unet/vision_transformer.py -- library GEMMs, reported as such).

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
        if self.version != 'V1' or split_gpu or do_LSA or do_SPT:
            raise NotImplementedError("b200unet.Generic_ViT_UNet: only vit_version='V1' on one device without "
                                      "LSA / SPT is implemented (no eager fallback)")
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
                                qkv_bias=True, task_specific_ln=ViT_task_specific_ln, task_name=first_task_name)
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
        if grad:
            skip0, token = _EncoderFunction.apply(self, plan, x, with_bott, *params)
        else:
            with torch.no_grad():
                skip0, token = _EncoderFunction.forward(_NullCtx(), self, plan, x, with_bott, *params)
        if store_vit_input:
            self.ViT_in = skip0.clone()
        if self.precision != "fp32" and self.ViT.native_supported(skip0):
            # bf16 mode: the whole ViT runs in the hand-written kernels of csrc/vit.cu; its input gradient is added straight
            # into the plan's gradient buffer of the first skip (so the encoder backward needs no extra add)
            bott = _view(plan, 2 * P + 1, 1)
            vit_out = self.ViT.forward_native(skip0, _view(plan, 1, 2) if grad else None, tuple(int(v) for v in bott.shape[1:]))
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
