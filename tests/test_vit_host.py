"""CPU: host-side mirror of the ViT-U-Net plugin surface (b200unet/generic_ViT_UNet.py, vision_transformer.py) --
constructor, module tree / state_dict keys / parameter order equal the reference's (pinned through oracle/vit_unet.py,
itself pinned to the reference's class in tests/test_oracle_vs_reference.py); unsupported variants raise instead of
falling back."""
import pytest
import torch

import util  # noqa: F401
from b200unet.generic_ViT_UNet import Generic_ViT_UNet, commDiv
from oracle import vit_unet


def _product(**kw):
    return Generic_ViT_UNet(1, 8, 3, 2, [16, 32, 32], pool_op_kernel_sizes=[[2, 2, 2], [2, 2, 2]],
                            conv_kernel_sizes=[[3, 3, 3]] * 3, **kw)


def test_tree_matches_reference_order():
    prod = _product()
    orc = vit_unet.Generic_ViT_UNet(1, 8, 3, 2, [16, 32, 32], [[2, 2, 2], [2, 2, 2]])
    assert [(n, tuple(p.shape)) for n, p in prod.named_parameters()] == [(n, tuple(p.shape)) for n, p in orc.named_parameters()]
    assert list(prod.state_dict().keys()) == list(orc.state_dict().keys())
    assert [n for n, _ in prod.named_children()] == ['conv_blocks_localization', 'conv_blocks_context', 'ViT', 'td', 'tu', 'seg_outputs']
    prod.load_state_dict(orc.state_dict())
    assert prod.patch_size == orc.patch_size == (16, 16) and prod.num_classesViT == orc.num_classesViT == 32 * 4 * 8 * 8
    assert prod.img_size == orc.img_size == [16, 32, 32] and prod.in_chans == 8
    assert prod.ViT.patch_embeds[0].num_patches == 4


def test_sizes_cfg4():
    # patch 16, 432 tokens, head 768 -> 320*6*6*6... (3 x 6 x 6 at cfg4's pooling)
    with torch.device("meta"):
        net = Generic_ViT_UNet(1, 32, 2, 5, [48, 192, 192], pool_op_kernel_sizes=[[2, 2, 2]] * 3 + [[1, 2, 2]] * 2,
                               conv_kernel_sizes=[[3, 3, 3]] * 6, weightInitializer=None)
    assert net.patch_size == (16, 16) and net.ViT.patch_embeds[0].num_patches == 432
    assert net.num_classesViT == 320 * 6 * 6 * 6
    n_vit = sum(p.numel() for p in net.ViT.parameters())
    assert 235e6 < n_vit < 245e6


def test_commdiv():
    assert commDiv(48, 192) == [1, 2, 3, 4, 6, 8, 12, 16, 24, 48]
    assert max(x for x in commDiv(20, 30) if x <= 16) == 10


@pytest.mark.parametrize("kw", [dict(vit_version='V4'), dict(do_SPT=True),
                                dict(split_gpu=True)])
def test_unsupported_variants_raise(kw):
    with pytest.raises(NotImplementedError):
        _product(**kw)


def test_cpu_forward_refuses():
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        _product()(torch.zeros(2, 1, 16, 32, 32))


def test_task_specific_layernorms_register_and_select():
    """vision_transformer.py:380-416: LNs live in ModuleDicts keyed by task; register_new_task adds a fresh set, use_task
    selects the set the forward (and the native kernel parameter table) uses"""
    net = _product(ViT_task_specific_ln=True, first_task_name='A')
    vit = net.ViT
    keys = list(vit.state_dict().keys())
    assert 'blocks.layer.0.norm1.A.weight' in keys and 'norm.A.bias' in keys and 'blocks.layer.11.norm2.A.weight' in keys
    with pytest.raises(AssertionError):
        vit._native_params()                 # no task selected yet (Block.forward asserts the same, :190)
    vit.use_task('A')
    n0 = len(list(vit.parameters()))
    vit.register_new_task('B')
    assert len(list(vit.parameters())) == n0 + 2 * (2 * 12 + 1)
    assert isinstance(vit.patch_embeds[0].norm['B'], torch.nn.Identity)
    with torch.no_grad():
        vit.blocks.layer[3].norm1['B'].weight.fill_(2.0)
    pa = vit._native_params()
    vit.use_task('B')
    pb = vit._native_params()
    assert len(pa) == len(pb) == 2 + 12 * 12 + 6
    i = 2 + 3 * 12
    assert pa[i] is vit.blocks.layer[3].norm1['A'].weight and pb[i] is vit.blocks.layer[3].norm1['B'].weight
    assert pb[-6] is vit.norm['B'].weight and pa[2 + 12 * 12] is vit.norm['A'].weight
    x = torch.randn(2, 8, 16, 32, 32)
    ya, yb = vit(x, task_name=None), None
    vit.use_task('A')
    yb = vit(x)
    assert not torch.allclose(ya, yb)        # the B set (norm1 weight 2) was used for ya
