"""ORACLE (test infrastructure, CPU/PyTorch fp32) -- restatement of the reference's ViT-U-Net hybrid, version V1 without
LSA / SPT / task-specific LayerNorms (the configuration `--use_vit` builds by default), so it can travel to the GPU box.

* VisionTransformer       -> reference nnunet_ext/network_architecture/vision_transformer.py:218-458 (ctor) and
                             :418-458 (forward_features / forward); PatchEmbed :16-79; Attention :120-151 (non-LSA
                             branch); Block :153-198; Encoder :200-216
* Generic_ViT_UNet        -> reference nnunet_ext/network_architecture/generic_ViT_UNet.py:21-214 (ctor: dry-run sizes
                             :85-131, patch_dim :148, ViT config :163-187, registration order :193-211) and :217-287
                             (forward), :290-296 (V1 input = first skip)
* timm pieces (un-vendored, timm@a41de1f) -> SURVEY.md Appendix A; restated separately in oracle/shim/timm so that the
  reference's own unmodified files can be imported HERE to pin this restatement (tests/test_oracle_vs_reference.py,
  tests/golden/vit_unet_tiny.npz).

state_dict keys and named_parameters() order equal the reference's, so weights move between the two by name.
Only tests/, __graft_entry__.smoke() and bench.py's CPU legs may import this module.
"""
import math
import os
import sys

import torch
from torch import nn

_SHIM = os.path.join(os.path.dirname(os.path.abspath(__file__)), "shim")
if _SHIM not in sys.path:
    sys.path.insert(0, _SHIM)

from nnunet.network_architecture.generic_UNet import Generic_UNet  # noqa: E402
from nnunet.network_architecture.initialization import InitWeights_He  # noqa: E402

VIT_TYPES = {'base': (768, 12, 12), 'large': (1024, 16, 24), 'huge': (1280, 16, 32)}  # embed, heads, layers (:64-66)


def common_divisors(a, b):
    """helpful_functions.py:272-286 (commDiv)."""
    n = math.gcd(a, b)
    return [i for i in range(1, n + 1) if n % i == 0]


class _Tokens3D(nn.Module):
    """3D patch embedding: cubic Conv3d with k = s = patch (vision_transformer.py:43-50), flatten, no norm (:75-78)."""

    def __init__(self, img_size, patch, in_chans, embed_dim):
        super().__init__()
        d, h, w = img_size
        self.num_patches = (w // patch) * (h // patch) * (d // patch)      # :47
        self.proj = nn.Conv3d(in_chans, embed_dim, kernel_size=patch, stride=patch)
        self.norm = nn.Identity()

    def forward(self, x):
        return self.norm(self.proj(x).flatten(2).transpose(1, 2))


class _Mlp(nn.Module):
    def __init__(self, dim, hidden):
        super().__init__()
        self.fc1, self.act, self.fc2 = nn.Linear(dim, hidden), nn.GELU(), nn.Linear(hidden, dim)

    def forward(self, x):
        return self.fc2(self.act(self.fc1(x)))


class _Attention(nn.Module):
    def __init__(self, dim, heads):
        super().__init__()
        self.num_heads, self.scale = heads, (dim // heads) ** -0.5
        self.qkv, self.proj = nn.Linear(dim, 3 * dim, bias=True), nn.Linear(dim, dim)

    def forward(self, x):   # :136-150
        B, N, C = x.shape
        q, k, v = self.qkv(x).reshape(B, N, 3, self.num_heads, C // self.num_heads).permute(2, 0, 3, 1, 4).unbind(0)
        w = ((q @ k.transpose(-2, -1)) * self.scale).softmax(dim=-1)
        return self.proj((w @ v).transpose(1, 2).reshape(B, N, C)), w


class _Block(nn.Module):
    def __init__(self, dim, heads, eps):
        super().__init__()
        self.norm1 = nn.LayerNorm(dim, eps=eps)
        self.attn = _Attention(dim, heads)
        self.norm2 = nn.LayerNorm(dim, eps=eps)
        self.mlp = _Mlp(dim, 4 * dim)

    def forward(self, x):   # :193-197
        a, w = self.attn(self.norm1(x))
        x = x + a
        return x + self.mlp(self.norm2(x)), w


class _Encoder(nn.Module):
    def __init__(self, depth, dim, heads, eps):
        super().__init__()
        self.layer = nn.ModuleList([_Block(dim, heads, eps) for _ in range(depth)])

    def forward(self, x):
        ws = []
        for blk in self.layer:
            x, w = blk(x)
            ws.append(w)
        return x, ws


class VisionTransformer(nn.Module):
    """3D ViT, one patch embedding / one head (the V1 build).  Parameter order = the reference's: cls_token,
    pos_embed_0, blocks.layer.*, norm, patch_embeds.0.proj, heads.0."""

    def __init__(self, img_size, patch, in_chans, num_classes, embed_dim, depth, heads):
        super().__init__()
        eps = 1e-6                                                        # norm_layer default (:226)
        self.cls_token = nn.Parameter(torch.zeros(1, 1, embed_dim))
        tokens = _Tokens3D(img_size, patch, in_chans, embed_dim)
        self.pos_embed_0 = nn.Parameter(torch.zeros(1, tokens.num_patches + 1, embed_dim))   # zeros (:364-366)
        self.blocks = _Encoder(depth, embed_dim, heads, eps)
        self.norm = nn.LayerNorm(embed_dim, eps=eps)
        self.patch_embeds = nn.ModuleList([tokens])
        self.heads = nn.ModuleList([nn.Linear(embed_dim, num_classes)])
        self.attn_weights = None
        nn.init.trunc_normal_(self.cls_token, std=.02)                   # timm init_weights('')
        nn.init.trunc_normal_(self.heads[0].weight, std=.02)
        nn.init.zeros_(self.heads[0].bias)

    def forward(self, x, idx=0, task_name=None):   # :418-458
        x = self.patch_embeds[idx](x)
        x = torch.cat((self.cls_token.expand(x.shape[0], -1, -1), x), dim=1) + getattr(self, 'pos_embed_%d' % idx)
        x, self.attn_weights = self.blocks(x)
        return self.heads[idx](self.norm(x)[:, 0])


class Generic_ViT_UNet(Generic_UNet):
    """V1: the ViT reads the first skip; its class-token head output, reshaped, replaces the bottleneck activation.
    V2 (:299-313): the ViT reads first skip + bottleneck up-sampled through all `tu`; V3 (:316-338): plus every skip up-sampled."""

    def __init__(self, input_channels, base_num_features, num_classes, num_pool, patch_size, pool_op_kernel_sizes,
                 conv_kernel_sizes=None, vit_type='base', max_num_features=None, vit_version='V1'):
        if conv_kernel_sizes is None:
            conv_kernel_sizes = [[3, 3, 3]] * (num_pool + 1)
        super().__init__(input_channels, base_num_features, num_classes, num_pool, 2, 2, nn.Conv3d, nn.InstanceNorm3d,
                         {'eps': 1e-5, 'affine': True}, nn.Dropout3d, {'p': 0, 'inplace': True}, nn.LeakyReLU,
                         {'negative_slope': 1e-2, 'inplace': True}, True, False, lambda x: x, InitWeights_He(1e-2),
                         pool_op_kernel_sizes, conv_kernel_sizes, False, True, True, max_num_features)
        # sizes of first skip and bottleneck (the reference finds them by a dry run, :85-131)
        with torch.no_grad():
            s = torch.zeros(1, input_channels, *patch_size)
            skip0 = None
            for d in range(len(self.conv_blocks_context) - 1):
                s = self.conv_blocks_context[d](s)
                skip0 = s.shape if skip0 is None else skip0
            bott = self.conv_blocks_context[-1](s).shape
        self.img_size = list(skip0[2:])
        self.num_classesViT = int(bott[1] * bott[2] * bott[3] * bott[4])
        p = max(x for x in common_divisors(self.img_size[0], self.img_size[1]) if x <= 16)   # :148
        self.patch_size, self.in_chans = (p, p), int(skip0[1])
        e, h, l = VIT_TYPES[vit_type.lower()]
        vit = VisionTransformer(self.img_size, p, self.in_chans, self.num_classesViT, e, l, h)
        # registration order of :193-211 (localization, context, ViT, td, tu, seg_outputs)
        parts = [(n, getattr(self, n)) for n in ('conv_blocks_localization', 'conv_blocks_context', 'td', 'tu', 'seg_outputs')]
        for n, _ in parts:
            delattr(self, n)
        table = dict(parts)
        for n in ('conv_blocks_localization', 'conv_blocks_context', 'ViT', 'td', 'tu', 'seg_outputs'):
            setattr(self, n, vit if n == 'ViT' else table[n])
        self.version = vit_version.title()
        assert self.version in ('V1', 'V2', 'V3')

    def vit_input(self, skips, last_context):   # :290-338
        if self.version == 'V1':
            return skips[0]
        t = last_context
        for u in range(len(self.tu)):
            t = self.tu[u](t)
        if self.version == 'V2':
            return skips[0] + t
        vit_in = torch.zeros(skips[0].size()) + t
        for idx, skip in enumerate(reversed(skips)):
            t = skip
            for u in range(idx + 1, len(self.tu)):
                t = self.tu[u](t)
            vit_in = vit_in + t
        return vit_in

    def forward(self, x):   # :217-287
        skips, seg = [], []
        for d in range(len(self.conv_blocks_context) - 1):
            x = self.conv_blocks_context[d](x)
            skips.append(x)
        x = self.conv_blocks_context[-1](x)          # V1: computed, only its size is used
        x = self.ViT(self.vit_input(skips, x)).reshape(x.size())
        for u in range(len(self.tu)):
            x = self.conv_blocks_localization[u](torch.cat((self.tu[u](x), skips[-(u + 1)]), dim=1))
            seg.append(self.final_nonlin(self.seg_outputs[u](x)))
        if self._deep_supervision and self.do_ds:
            return tuple([seg[-1]] + list(seg[:-1][::-1]))
        return seg[-1]


def fill_parameters(net, seed=0):
    """Deterministic, init-independent weights (name-order, one generator): lets the reference class, this restatement
    and the CUDA build hold identical parameters without shipping 10^8 numbers."""
    g = torch.Generator().manual_seed(seed)
    with torch.no_grad():
        for name, p in net.named_parameters():
            if name.endswith('instnorm.weight') or ('norm' in name and name.endswith('.weight') and p.dim() == 1):
                p.copy_(1.0 + 0.1 * torch.randn(p.shape, generator=g))
            elif p.dim() == 1 or name.endswith('bias'):
                p.copy_(0.05 * torch.randn(p.shape, generator=g))
            elif 'ViT' in name:
                p.copy_(0.02 * torch.randn(p.shape, generator=g))
            else:
                fan_in = p[0].numel()
                p.copy_(torch.randn(p.shape, generator=g) * (2.0 / fan_in) ** 0.5)
    return net
