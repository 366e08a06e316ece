"""CPU, build container only: pins the oracle restatements (oracle/cl_losses.py) to the REFERENCE'S OWN unmodified
files, imported from /root/reference with the nnunet shim on sys.path, and re-checks the committed golden fixtures.
Skipped where the reference is not mounted (the GPU box)."""
import os
import sys

import numpy as np
import pytest
import torch

import util

REF = "/root/reference"
pytestmark = pytest.mark.skipif(not os.path.isdir(REF), reason="reference not mounted")


@pytest.fixture(scope="module")
def refvals():
    for p in (os.path.join(util.ROOT, "oracle", "shim"), REF):
        if p not in sys.path:
            sys.path.insert(0, p)
    from oracle import gen_golden
    case = gen_golden.seeded_case()
    return case, gen_golden.reference_values(case)


def test_golden_fixture_is_current(refvals):
    """tests/golden/cl_losses.npz == what the reference produces today"""
    _, vals = refvals
    gold = np.load(os.path.join(util.ROOT, "tests", "golden", "cl_losses.npz"))
    assert set(gold.files) == set(vals.keys())
    for k in gold.files:
        np.testing.assert_allclose(gold[k], vals[k], rtol=1e-6, atol=1e-7, equal_nan=True, err_msg=k)


def test_reference_own_structural_test_passes_on_the_shim():
    """the reference's test/network_architecture/test_MultiHead_Module.py runs unmodified against the shim"""
    import subprocess
    env = dict(os.environ, PYTHONPATH=os.pathsep.join([os.path.join(util.ROOT, "oracle", "shim"), REF]))
    r = subprocess.run([sys.executable, "-m", "pytest", "-x", "-q", "-p", "no:cacheprovider",
                        os.path.join(REF, "test", "network_architecture", "test_MultiHead_Module.py")],
                       env=env, capture_output=True, text=True, cwd="/tmp")
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]


def test_vit_unet_restatement_equals_reference_class():
    """oracle/vit_unet.py == the reference's generic_ViT_UNet.py + vision_transformer.py (unmodified, on the nnunet /
    timm shims): same state_dict keys and parameter order, identical logits and gradients, and the committed fixture
    is what the reference produces today."""
    for p in (os.path.join(util.ROOT, "oracle", "shim"), REF):
        if p not in sys.path:
            sys.path.insert(0, p)
    from oracle import gen_golden, vit_unet
    ref = gen_golden.reference_vit_unet()
    mine = vit_unet.Generic_ViT_UNet(1, 8, 3, 2, [16, 32, 32], [[2, 2, 2], [2, 2, 2]])
    assert [(n, tuple(q.shape)) for n, q in ref.named_parameters()] == [(n, tuple(q.shape)) for n, q in mine.named_parameters()]
    assert list(ref.state_dict().keys()) == list(mine.state_dict().keys())
    mine.load_state_dict(ref.state_dict())
    a, b = gen_golden.vit_values(ref), gen_golden.vit_values(mine)
    assert set(a) == set(b)
    for k in a:
        np.testing.assert_allclose(a[k], b[k], rtol=1e-6, atol=1e-7, err_msg=k)
    assert a["gnorm/conv_blocks_context.2.0.blocks.0.conv.weight"] == -1.0     # Q14: bottleneck gets no gradient
    gold = np.load(os.path.join(util.ROOT, "tests", "golden", "vit_unet_tiny.npz"))
    assert set(gold.files) == set(a)
    for k in gold.files:
        np.testing.assert_allclose(gold[k], a[k], rtol=1e-5, atol=1e-6, err_msg=k)


@pytest.mark.parametrize("n_finished", [1, 2, 3])
def test_rw_finish_task_restatement(n_finished):
    """oracle.cl_losses.rw_finish_task == the reference's OWN lines rw/nnUNetTrainerRW.py:180-200, extracted from the
    unmodified file and executed against a stand-in `self` (the trainer class itself needs nnunet's trainer stack)."""
    import textwrap
    import types
    from oracle import cl_losses
    src = open(os.path.join(REF, "nnunet_ext/training/network_training/rw/nnUNetTrainerRW.py")).read().splitlines()
    lo = next(i for i, l in enumerate(src) if "Normalize the fisher values to be in range" in l)
    hi = next(i for i, l in enumerate(src) if "Store the fisher and param values" in l and i > lo)
    block = textwrap.dedent("\n".join(src[lo:hi]))
    g = torch.Generator().manual_seed(3)
    names = ["a.weight", "a.bias", "b.weight"]
    shapes = [(4, 3, 3), (4,), (2, 5)]
    tasks = ["A", "B", "C"][:n_finished]
    cur = tasks[-1]
    fisher = {t: {n: torch.rand(s, generator=g) for n, s in zip(names, shapes)} for t in tasks}
    scores = {t: {n: 3 * torch.rand(s, generator=g) for n, s in zip(names, shapes)} for t in tasks}
    me_f, me_s = cl_losses.rw_finish_task({k: v.clone() for k, v in fisher[cur].items()},
                                          {k: v.clone() for k, v in scores[cur].items()}, n_finished)
    self = types.SimpleNamespace(task=cur, fold=0, fisher=fisher, scores=scores,
                                 already_trained_on={"0": {"finished_training_on": list(tasks)}})
    exec(block, {"torch": torch, "EPSILON": 1e-8, "self": self})
    for n in names:
        assert torch.equal(self.fisher[cur][n], me_f[n]), n
        assert torch.equal(self.scores[cur][n], me_s[n]), n


def test_reference_own_multihead_test_passes_on_b200unet_multihead_module(tmp_path):
    """the reference's test/network_architecture/test_MultiHead_Module.py, unmodified, with
    ``nnunet_ext.network_architecture.MultiHead_Module`` resolved to b200unet's from-scratch mirror (nnunet_ext is a
    namespace package: a directory earlier on sys.path that holds only that one module shadows the reference's)."""
    import subprocess
    d = tmp_path / "nnunet_ext" / "network_architecture"
    d.mkdir(parents=True)
    (d / "MultiHead_Module.py").write_text("from b200unet.MultiHead_Module import *  # noqa\n"
                                           "from b200unet.MultiHead_Module import MultiHead_Module  # noqa\n")
    env = dict(os.environ, PYTHONPATH=os.pathsep.join([str(tmp_path), util.PKG, os.path.join(util.ROOT, "oracle", "shim"), REF]))
    r = subprocess.run([sys.executable, "-m", "pytest", "-x", "-q", "-p", "no:cacheprovider",
                        os.path.join(REF, "test", "network_architecture", "test_MultiHead_Module.py")],
                       env=env, capture_output=True, text=True, cwd="/tmp")
    assert r.returncode == 0, r.stdout[-3000:] + r.stderr[-2000:]


def test_task_specific_ln_vit_equals_reference_class():
    """b200unet.VisionTransformer with task-specific LayerNorms == the reference's class (vision_transformer.py, unmodified,
    on the timm shim): state_dict keys / parameter order after register_new_task, identical outputs per selected task"""
    for p in (os.path.join(util.ROOT, "oracle", "shim"), REF):
        if p not in sys.path:
            sys.path.insert(0, p)
    from nnunet_ext.network_architecture.vision_transformer import PatchEmbed as RefPE, VisionTransformer as RefViT
    from b200unet.vision_transformer import VisionTransformer
    kw = dict(ViT_2d=False, img_size=[16, 32, 32], patch_size=(16, 16), img_depth=[16], in_chans=8, num_classes=96,
              embed_dim=128, depth=2, num_heads=2, mlp_ratio=4, qkv_bias=True, task_specific_ln=True, task_name='A')
    torch.manual_seed(0)
    ref = RefViT(representation_size=None, distilled=False, drop_rate=0, attn_drop_rate=0, drop_path_rate=0, embed_layer=RefPE,
                 norm_layer=None, act_layer=None, weight_init='', **kw)
    mine = VisionTransformer(**kw)
    for v in (ref, mine):
        v.register_new_task('B')
    assert [(n, tuple(q.shape)) for n, q in ref.named_parameters()] == [(n, tuple(q.shape)) for n, q in mine.named_parameters()]
    assert list(ref.state_dict().keys()) == list(mine.state_dict().keys())
    g = torch.Generator().manual_seed(3)
    with torch.no_grad():
        for n, q in ref.named_parameters():
            if 'norm' in n or 'pos_embed' in n:
                q.copy_(1 + 0.2 * torch.randn(q.shape, generator=g))
    mine.load_state_dict(ref.state_dict())
    x = torch.randn((2, 8, 16, 32, 32), generator=g)
    for task in ('A', 'B'):
        ref.use_task(task)
        mine.use_task(task)
        np.testing.assert_allclose(mine(x).detach().numpy(), ref(x).detach().numpy(), rtol=1e-5, atol=1e-6)
    a = mine(x, task_name='A')
    mine.use_task('B')
    assert not torch.allclose(a, mine(x))


@pytest.mark.parametrize("version", ["V2", "V3"])
def test_vit_unet_v2_v3_restatement_equals_reference_class(version):
    """oracle/vit_unet.py V2 / V3 (ViT input fused from the first skip, the up-sampled bottleneck and -- V3 -- every up-sampled
    skip) == the reference's class (generic_ViT_UNet.py:290-338): identical logits and gradient norms, and here the bottleneck
    convolutions DO receive gradients."""
    for p in (os.path.join(util.ROOT, "oracle", "shim"), REF):
        if p not in sys.path:
            sys.path.insert(0, p)
    from oracle import gen_golden, vit_unet
    ref = gen_golden.reference_vit_unet(version)
    mine = vit_unet.Generic_ViT_UNet(1, 8, 3, 2, [16, 32, 32], [[2, 2, 2], [2, 2, 2]], vit_version=version)
    assert list(ref.state_dict().keys()) == list(mine.state_dict().keys())
    mine.load_state_dict(ref.state_dict())
    a, b = gen_golden.vit_values(ref), gen_golden.vit_values(mine)
    assert set(a) == set(b)
    for k in a:
        np.testing.assert_allclose(a[k], b[k], rtol=1e-6, atol=1e-7, err_msg=k)
    assert a["gnorm/conv_blocks_context.2.0.blocks.0.conv.weight"] > 0


def test_lsa_vit_equals_reference_class():
    """b200unet.VisionTransformer with Locality Self-Attention == the reference's class (vision_transformer.py:81-151): parameter
    names / order (attn.scale first, bias-free qkv, unused timm proj, to_out), state_dict keys, outputs and gradients"""
    for p in (os.path.join(util.ROOT, "oracle", "shim"), REF):
        if p not in sys.path:
            sys.path.insert(0, p)
    from nnunet_ext.network_architecture.vision_transformer import PatchEmbed as RefPE, VisionTransformer as RefViT
    from b200unet.vision_transformer import VisionTransformer
    kw = dict(ViT_2d=False, img_size=[16, 32, 32], patch_size=(8, 8), img_depth=[16], in_chans=8, num_classes=96,
              embed_dim=128, depth=2, num_heads=2, mlp_ratio=4, qkv_bias=True, is_LSA=True)
    torch.manual_seed(0)
    ref = RefViT(representation_size=None, distilled=False, drop_rate=0, attn_drop_rate=0, drop_path_rate=0, embed_layer=RefPE,
                 norm_layer=None, act_layer=None, weight_init='', **kw)
    mine = VisionTransformer(**kw)
    assert [(n, tuple(q.shape)) for n, q in ref.named_parameters()] == [(n, tuple(q.shape)) for n, q in mine.named_parameters()]
    assert list(ref.state_dict().keys()) == list(mine.state_dict().keys())
    g = torch.Generator().manual_seed(3)
    with torch.no_grad():
        for n, q in ref.named_parameters():
            if 'scale' in n or 'pos_embed' in n or 'norm' in n:
                q.add_(0.2 * torch.randn(q.shape, generator=g))
    mine.load_state_dict(ref.state_dict())
    x = torch.randn((2, 8, 16, 32, 32), generator=g)
    u = torch.randn((2, 96), generator=g)
    (ref(x) * u).sum().backward()
    (mine(x) * u).sum().backward()
    np.testing.assert_allclose(mine(x).detach().numpy(), ref(x).detach().numpy(), rtol=1e-5, atol=1e-5)
    rg = dict(ref.named_parameters())
    for n, q in mine.named_parameters():
        if rg[n].grad is None:
            assert q.grad is None, n            # timm's proj is registered but unused under LSA
        else:
            np.testing.assert_allclose(q.grad.numpy(), rg[n].grad.numpy(), rtol=1e-4, atol=2e-5, err_msg=n)
    assert rg["blocks.layer.0.attn.scale"].grad is not None and rg["blocks.layer.0.attn.proj.weight"].grad is None
