"""``VisionTransformer`` -- host-side mirror of the reference's 3D ViT (reference
nnunet_ext/network_architecture/vision_transformer.py:218-458) for the ``Generic_ViT_UNet`` V1 build: one 3D patch
embedding, one head, optionally task-specific LayerNorms (:380-416); LSA / SPT raise (SPT cannot run on 3D inputs in the
reference either: its Rearrange pattern is 4-D).

Two execution paths:
* bf16 mode (production): the whole ViT -- patch embedding, 12 blocks, final norm, head, forward AND backward -- runs in
  hand-written kernels behind the C ABI (csrc/vit.cu, b2_vit_forward / b2_vit_backward): every Linear on the tcgen05
  gather-GEMM, a shared-memory attention kernel with recomputation in backward, LayerNorm / GELU / reductions as fused
  HBM-bound kernels.  No ATen / cuBLAS kernel is launched.  The first skip is consumed in place as a channels-last view of
  the U-Net plan's workspace and the input gradient is added straight into the plan's gradient buffer.
* fp32 parity mode (and `store_attn_weights`): the same arithmetic through ATen in fp32 (library GEMMs), the reference
  the native path is tested against at the 1e-3 level of the oracle.

state_dict keys / named_parameters() order are the reference's (cls_token, pos_embed_0, blocks.layer.N.{norm1,
attn.{qkv,proj},norm2,mlp.{fc1,fc2}}, norm, patch_embeds.0.proj, heads.0) so checkpoints and name-keyed Fisher
dictionaries (`'ViT'` / `'norm'` substring filters, ewc_ln/nnUNetTrainerEWCLN.py:49-50) keep working.
"""
import ctypes as C

import torch
import torch.nn.functional as F
from torch import nn

from . import _lib

VIT_TYPES = {'base': {'embed_size': 768, 'head': 12, 'layers': 12},
             'large': {'embed_size': 1024, 'head': 16, 'layers': 24},
             'huge': {'embed_size': 1280, 'head': 16, 'layers': 32}}   # generic_ViT_UNet.py:64-66


class PatchEmbed(nn.Module):
    """vision_transformer.py:16-79, 3D branch: cubic patches of edge `patch`, flattened in (d, h, w) raster order."""

    def __init__(self, img_size, patch, in_chans, embed_dim, task_specific_ln=False, task_name=None):
        super().__init__()
        d, h, w = (int(s) for s in img_size)
        self.img_size, self.patch_size = (d, h, w), (d, patch, patch)     # attribute quirk of :43-44 kept
        self.patch = int(patch)
        self.grid_size = (d // patch, h // patch, w // patch)
        self.num_patches = (w // patch) * (h // patch) * (d // patch)     # :47
        self.flatten, self.embed2D, self.task_specific_ln = True, False, bool(task_specific_ln)
        self.proj = nn.Conv3d(in_chans, embed_dim, kernel_size=patch, stride=patch)   # parameter container
        # Generic_ViT_UNet passes norm_layer=None: Identity, or a ModuleDict of Identities per task (:55-57)
        self.norm = nn.ModuleDict({task_name: nn.Identity()}) if self.task_specific_ln else nn.Identity()

    def forward(self, x, task_name=None):
        B, Cc, D, H, W = x.shape
        if H != self.img_size[1] or W != self.img_size[2]:
            raise AssertionError("Input image size (%d,%d) doesn't match model (%d,%d)." % (H, W, self.img_size[1], self.img_size[2]))
        p = self.patch
        gd, gh, gw = D // p, H // p, W // p
        if (gd * p, gh * p, gw * p) != (D, H, W):
            x = x[:, :, :gd * p, :gh * p, :gw * p]                         # Conv3d floor semantics
        tok = x.reshape(B, Cc, gd, p, gh, p, gw, p).permute(0, 2, 4, 6, 1, 3, 5, 7).reshape(B, gd * gh * gw, Cc * p ** 3)
        out = F.linear(tok, self.proj.weight.reshape(self.proj.out_channels, -1), self.proj.bias)
        if self.proj._forward_hooks:
            # PLOP / POD hook every 'conv.Conv*' module (reference plop:330-353), the patch-embedding Conv3d included: hand
            # the hooks what that module returns, (B, E, D/p, H/p, W/p)
            conv_out = out.transpose(1, 2).reshape(B, -1, gd, gh, gw)
            for hook in list(self.proj._forward_hooks.values()):
                hook(self.proj, (x,), conv_out)
        return out


class Mlp(nn.Module):
    def __init__(self, dim, hidden):
        super().__init__()
        self.fc1 = nn.Linear(dim, hidden)
        self.act = nn.GELU()
        self.fc2 = nn.Linear(hidden, dim)

    def forward(self, x):
        return self.fc2(self.act(self.fc1(x)))


class Attention(nn.Module):
    """vision_transformer.py:81-151.  Plain branch (:136-150): qkv with bias, fixed scale head_dim^-0.5, `proj`.  LSA branch
    (Locality Self-Attention, :90-135): qkv WITHOUT bias (xavier-normal), a learnable temperature per head (`scale`), the diagonal
    of the score matrix masked out (-987654321 before the softmax), output through `to_out`; timm's `proj` stays registered but
    unused, as in the reference.  The probabilities are only materialised when the owner asked for them
    (``VisionTransformer.store_attn_weights``) or for LSA in the ATen path; otherwise the fused SDPA kernel is used."""

    def __init__(self, dim, num_heads, is_LSA=False, num_patches=16):
        super().__init__()
        self.num_heads = num_heads
        self.LSA = bool(is_LSA)
        self.qkv = nn.Linear(dim, dim * 3, bias=not self.LSA)
        self.proj = nn.Linear(dim, dim)
        self.store_weights = False
        if self.LSA:
            self.num_patches, self.heads, self.dim, self.inner_dim = num_patches, num_heads, dim, (dim // num_heads) * num_heads
            nn.init.xavier_normal_(self.qkv.weight)
            self.to_out = nn.Sequential(nn.Linear(self.inner_dim, dim), nn.Dropout(0.))
            self.scale = nn.Parameter((dim // num_heads) ** -0.5 * torch.ones(num_heads))
            self.mask = torch.nonzero(torch.eye(num_patches + 1, num_patches + 1) == 1, as_tuple=False)
        else:
            self.scale = (dim // num_heads) ** -0.5

    def forward(self, x):
        B, N, Cc = x.shape
        q, k, v = self.qkv(x).reshape(B, N, 3, self.num_heads, Cc // self.num_heads).permute(2, 0, 3, 1, 4).unbind(0)
        if self.LSA:
            dots = (q @ k.transpose(-2, -1)) * self.scale.to(q.dtype).view(1, -1, 1, 1)
            # quirk: the reference sizes the mask from the 2-D patch count (vision_transformer.py:289-294), so in 3-D only the
            # first num_patches + 1 tokens mask their own score
            m = torch.zeros(N, dtype=torch.bool, device=x.device)
            m[:self.num_patches + 1] = True
            dots = dots.masked_fill(torch.diag(m), -987654321)
            w = dots.softmax(dim=-1)
            return self.to_out((w @ v).transpose(1, 2).reshape(B, N, Cc)), w
        if self.store_weights:
            w = ((q @ k.transpose(-2, -1)) * self.scale).softmax(dim=-1)
            o = w @ v
        else:
            w = None
            o = F.scaled_dot_product_attention(q, k, v, scale=self.scale)
        return self.proj(o.transpose(1, 2).reshape(B, N, Cc)), w


class Block(nn.Module):
    """vision_transformer.py:153-198: with task-specific LNs, norm1 / norm2 are ModuleDicts keyed by task name and
    `use_task_name` (set through VisionTransformer.use_task) selects the pair used by forward"""

    def __init__(self, dim, num_heads, eps, task_specific_ln=False, task_name=None, is_LSA=False, num_patches=16):
        super().__init__()
        self.task_specific_ln = bool(task_specific_ln)
        if self.task_specific_ln:
            self.use_task_name = None
            assert task_name is not None and isinstance(task_name, str), \
                "When using task specific LNs, than please provide a task_name during initialization.."
        mk = lambda: nn.ModuleDict({task_name: nn.LayerNorm(dim, eps=eps)}) if self.task_specific_ln else nn.LayerNorm(dim, eps=eps)
        self.norm1 = mk()
        self.attn = Attention(dim, num_heads, is_LSA, num_patches)
        self.drop_path = nn.Identity()
        self.norm2 = mk()
        self.mlp = Mlp(dim, 4 * dim)

    def norms(self):
        """the (norm1, norm2) LayerNorm pair the next forward uses"""
        if not self.task_specific_ln:
            return self.norm1, self.norm2
        assert self.use_task_name is not None and isinstance(self.use_task_name, str), \
            "When using task specific LNs, than please set a task_name for the forward call using ViT.use_task(..).."
        return self.norm1[self.use_task_name], self.norm2[self.use_task_name]

    def forward(self, x):
        n1, n2 = self.norms()
        a, w = self.attn(n1(x))
        x = x + a
        return x + self.mlp(n2(x)), w


class Encoder(nn.Module):
    def __init__(self, depth, dim, num_heads, eps, task_specific_ln=False, task_name=None, is_LSA=False, num_patches=16):
        super().__init__()
        self.layer = nn.ModuleList([Block(dim, num_heads, eps, task_specific_ln, task_name, is_LSA, num_patches) for _ in range(depth)])

    def forward(self, x):
        ws = []
        for blk in self.layer:
            x, w = blk(x)
            ws.append(w)
        return x, ws


class VisionTransformer(nn.Module):
    def __init__(self, ViT_2d, img_size, patch_size, img_depth, in_chans, num_classes, embed_dim=768, depth=12,
                 num_heads=12, mlp_ratio=4, qkv_bias=True, task_specific_ln=False, task_name=None, is_LSA=False,
                 is_SPT=False, **ignored):
        super().__init__()
        if task_specific_ln:
            assert not is_SPT and not is_LSA, "Currently, we do not provide the combination for task specific LNs and either LSA, SPT or both.."
            assert task_name is not None and isinstance(task_name, str), \
                "When using task specific LNs, than please provide a task_name during initialization.."
        if ViT_2d or is_SPT or mlp_ratio != 4 or not qkv_bias:
            raise NotImplementedError("b200unet.VisionTransformer: only the 3D build without SPT is implemented (SPT is 2-D only "
                                      "in the reference: its Rearrange pattern takes 4-D inputs)")
        self.LSA, self.SPT, self.task_specific_ln = bool(is_LSA), False, bool(task_specific_ln)
        self.task_name_use = None
        self.block_depth, self.embed_dim, self.num_features = depth, embed_dim, embed_dim
        self.num_tokens, self.num_classes = 1, int(num_classes)
        self.attn_weights = None
        self.store_attn_weights = False
        eps = 1e-6
        self.cls_token = nn.Parameter(torch.zeros(1, 1, embed_dim))
        pe = PatchEmbed((img_depth[0], img_size[1], img_size[2]) if len(img_size) == 3 else (img_depth[0],) + tuple(img_size),
                        patch_size[0], in_chans, embed_dim, self.task_specific_ln, task_name)
        self.pos_embed_0 = nn.Parameter(torch.zeros(1, pe.num_patches + 1, embed_dim))
        self.pos_drop = nn.Identity()
        # (the reference rebuilds the blocks WITHOUT LSA when task-specific LNs are on, :303-306 -- the combination is asserted away)
        # LSA's mask size comes from the 2-D timm patch embedding that exists at that point of the reference's constructor (:289)
        n2d = (pe.img_size[1] // pe.patch) * (pe.img_size[2] // pe.patch)
        self.blocks = Encoder(depth, embed_dim, num_heads, eps, self.task_specific_ln, task_name, self.LSA, n2d)
        self._ln_eps = eps
        self.norm = nn.ModuleDict({task_name: nn.LayerNorm(embed_dim, eps=eps)}) if self.task_specific_ln \
            else nn.LayerNorm(embed_dim, eps=eps)
        self.pre_logits = nn.Identity()
        self.patch_embeds = nn.ModuleList([pe])
        self.heads = nn.ModuleList([nn.Linear(embed_dim, self.num_classes)])
        self.head_dists = None
        nn.init.trunc_normal_(self.cls_token, std=.02)
        nn.init.trunc_normal_(self.heads[0].weight, std=.02)
        nn.init.zeros_(self.heads[0].bias)

    # -- native path (csrc/vit.cu) ------------------------------------------------------------------------------------------
    def _native_params(self):
        """parameters in named_parameters() order of the reference module (what b2_vit_forward expects)"""
        ps = [self.cls_token, self.pos_embed_0]
        for blk in self.blocks.layer:
            n1, n2 = blk.norms()          # task-specific LNs: the active task's pair (the kernels see plain LayerNorms)
            a = blk.attn
            out = a.to_out[0] if a.LSA else a.proj      # LSA: no qkv bias (None -> null pointer), output through `to_out`
            ps += [n1.weight, n1.bias, a.qkv.weight, a.qkv.bias, out.weight, out.bias,
                   n2.weight, n2.bias, blk.mlp.fc1.weight, blk.mlp.fc1.bias, blk.mlp.fc2.weight, blk.mlp.fc2.bias]
        pe = self.patch_embeds[0]
        norm = self._final_norm(None)
        ps += [norm.weight, norm.bias, pe.proj.weight, pe.proj.bias, self.heads[0].weight, self.heads[0].bias]
        if self.LSA:                      # one learnable temperature vector [heads] per block, after the tail
            ps += [blk.attn.scale for blk in self.blocks.layer]
        return ps

    def _final_norm(self, task_name):
        if not self.task_specific_ln:
            return self.norm
        if task_name is None:
            assert self.task_name_use is not None, ("Either set the task_name during forward or using the use_task function "
                                                    "when training with task specific LNs..")
            task_name = self.task_name_use
        return self.norm[task_name]

    def native_supported(self, x):
        blk = self.blocks.layer[0]
        return (x.is_cuda and x.dtype == torch.bfloat16 and not self.store_attn_weights and self.embed_dim == blk.attn.num_heads * 64
                and self.embed_dim <= 1024 and x.shape[0] <= 4 and x.shape[1] % 8 == 0 and x.stride(1) == 1)

    def _native_plan(self, x, out_shape):
        key = (tuple(x.shape), x.device, tuple(out_shape))
        plans = self.__dict__.setdefault('_vplans', {})
        if key not in plans:
            lib = _lib.load()
            d = _lib.VitDesc()
            d.batch, d.in_channels, d.D, d.H, d.W = (int(v) for v in x.shape)
            d.patch = self.patch_embeds[0].patch
            d.embed, d.heads, d.depth, d.mlp_ratio = self.embed_dim, self.blocks.layer[0].attn.num_heads, self.block_depth, 4
            d.out_features = self.num_classes
            d.out_c, d.out_d, d.out_h, d.out_w = (int(v) for v in out_shape)
            d.ln_eps = float(self._ln_eps)
            d.lsa = int(self.LSA)
            d.lsa_mask = int(self.blocks.layer[0].attn.num_patches + 1) if self.LSA else 0
            h = C.c_void_p()
            _lib.check(lib.b2_vit_plan_create(C.byref(d), C.byref(h)))
            ws = torch.empty(int(lib.b2_vit_workspace_bytes(h)), dtype=torch.uint8, device=x.device)
            plans[key] = (h, ws)
        return plans[key]

    def __getstate__(self):
        d = self.__dict__.copy()
        d.pop('_vplans', None)          # C handles / workspaces are never copied (teachers are deep copies of the network)
        return d

    def forward_native(self, x, dskip, out_shape, return_input_grad=False):
        """x: channels-last bf16 tensor (B, C, D, H, W), e.g. the first skip as a view of the U-Net plan; dskip: the matching
        gradient view of the plan (the input gradient is ADDED into it by the backward kernels) or None; return_input_grad:
        hand the input gradient to autograd instead (V2 / V3, where the ViT input is computed by further ops); returns the
        dense [B, F] fp32 head output"""
        return _ViTNativeFunction.apply(self, x, dskip, tuple(out_shape), bool(return_input_grad), *self._native_params())

    def register_new_task(self, task_name):
        """vision_transformer.py:380-401: a fresh LayerNorm set (final norm, every block's norm1 / norm2, Identity for the
        patch embedding) under `task_name`; selecting it is use_task's job"""
        assert not (self.SPT or self.LSA), "When using SPT or LSA, task specific LNs are not allowed, so you can not call this function.."
        assert self.task_specific_ln, "register_new_task needs a ViT built with task_specific_ln=True"
        ref = next(iter(self.norm.values())).weight
        mk = lambda: nn.LayerNorm(self.embed_dim, eps=self._ln_eps).to(device=ref.device, dtype=ref.dtype)
        self.norm[task_name] = mk()
        for pe in self.patch_embeds:
            pe.norm[task_name] = nn.Identity()
        for blk in self.blocks.layer:
            blk.norm1[task_name], blk.norm2[task_name] = mk(), mk()

    def use_task(self, task_name):
        """vision_transformer.py:403-416"""
        assert not (self.SPT or self.LSA), "When using SPT or LSA, task specific LNs are not allowed, so you can not call this function.."
        self.task_name_use = task_name
        for blk in self.blocks.layer:
            blk.use_task_name = task_name

    def forward(self, x, idx=0, task_name=None):   # :418-458
        for blk in self.blocks.layer:
            blk.attn.store_weights = self.store_attn_weights
        norm = self._final_norm(task_name)
        x = self.patch_embeds[idx](x, task_name)
        cls = self.cls_token.expand(x.shape[0], -1, -1).to(x.dtype)
        x = torch.cat((cls, x), dim=1) + getattr(self, 'pos_embed_' + str(idx)).to(x.dtype)
        x, ws = self.blocks(x)
        self.attn_weights = ws if self.store_attn_weights else None
        return self.heads[idx](self.pre_logits(norm(x)[:, 0]))


def _view_of(t):
    from .deep_supervision import _act_view
    v, keep = _act_view(t)
    if keep.data_ptr() != t.data_ptr():
        raise RuntimeError("ViT native path needs channels-last bf16 views of the plan's workspace")
    return v


class _ViTNativeFunction(torch.autograd.Function):
    @staticmethod
    def forward(ctx, vit, x, dskip, out_shape, return_dx, *params):
        lib = _lib.load()
        h, ws = vit._native_plan(x, out_shape)
        for p in params:
            if p is not None and (p.dtype != torch.float32 or not p.is_contiguous() or p.device != x.device):
                raise RuntimeError("ViT parameters must be contiguous fp32 tensors on the input's device")
        need = any(ctx.needs_input_grad)      # (grad mode is off inside Function.forward: is_grad_enabled() would say False)
        pp = (C.c_void_p * len(params))(*[None if p is None else p.data_ptr() for p in params])
        out = torch.empty((x.shape[0], vit.num_classes), dtype=torch.float32, device=x.device)
        vx = _view_of(x.detach())
        _lib.check(lib.b2_vit_forward(h, pp, C.byref(vx), C.c_void_p(ws.data_ptr()), None, C.c_void_p(out.data_ptr()), 1 if need else 0,
                                      C.c_void_p(torch.cuda.current_stream(x.device).cuda_stream)))
        pe = vit.patch_embeds[0]
        if pe.proj._forward_hooks:      # PLOP / POD hook every conv module (reference plop:330-353): the patch-embedding output
            gd, gh, gw = (int(x.shape[2 + i]) // pe.patch for i in range(3))
            n_tok = x.shape[0] * gd * gh * gw
            off = int(lib.b2_vit_tokens_offset(h))
            tok = ws[off:off + n_tok * vit.embed_dim * 2].view(torch.bfloat16)
            conv_out = tok.view(x.shape[0], gd, gh, gw, vit.embed_dim).permute(0, 4, 1, 2, 3)
            for hook in list(pe.proj._forward_hooks.values()):
                hook(pe.proj, (x,), conv_out)
        ctx.vit, ctx.h, ctx.ws, ctx.params, ctx.dskip, ctx.x, ctx.return_dx = vit, h, ws, params, dskip, x, return_dx
        ctx.set_materialize_grads(False)
        return out

    @staticmethod
    def backward(ctx, dout):
        lib = _lib.load()
        params = ctx.params
        dev = dout.device
        dout = dout.contiguous().float()
        total = sum(p.numel() for p in params if p is not None)
        flat = torch.empty(total, dtype=torch.float32, device=dev)
        grads, o = [], 0
        for p in params:
            if p is None:
                grads.append(None)
                continue
            grads.append(flat[o:o + p.numel()].view(p.shape))
            o += p.numel()
        pp = (C.c_void_p * len(params))(*[None if p is None else p.data_ptr() for p in params])
        gp = (C.c_void_p * len(params))(*[None if g is None else g.data_ptr() for g in grads])
        dx = None
        if ctx.return_dx:          # the kernels ADD the input gradient into the buffer they are given
            dx = torch.empty(tuple(ctx.x.shape), dtype=ctx.x.dtype, device=dev, memory_format=torch.channels_last_3d).zero_()
            vd = _view_of(dx)
        else:
            vd = _view_of(ctx.dskip) if ctx.dskip is not None else None
        _lib.check(lib.b2_vit_backward(ctx.h, pp, None, C.c_void_p(dout.data_ptr()), C.c_void_p(ctx.ws.data_ptr()),
                                       None if vd is None else C.byref(vd), gp,
                                       C.c_void_p(torch.cuda.current_stream(dev).cuda_stream)))
        ctx.vit._last_flat_grad = flat
        return (None, dx, None, None, None) + tuple(g if (p is not None and p.requires_grad) else None for g, p in zip(grads, params))
