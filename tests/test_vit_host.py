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


@pytest.mark.parametrize("kw", [dict(vit_version='V2'), dict(vit_version='V4'), dict(do_LSA=True), dict(do_SPT=True),
                                dict(ViT_task_specific_ln=True, first_task_name='a'), dict(split_gpu=True)])
def test_unsupported_variants_raise(kw):
    with pytest.raises(NotImplementedError):
        _product(**kw)


def test_cpu_forward_refuses():
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        _product()(torch.zeros(2, 1, 16, 32, 32))
