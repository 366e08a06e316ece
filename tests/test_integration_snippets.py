"""CPU: the binding snippets of INTEGRATION.md are executed, not just printed.  (a) the ctypes stub against the built library;
(b) / (c) the name rebinding against the reference's own modules where they import here (reference mounted + nnunet / timm shims),
with empty stand-ins for the trainer modules whose nnunet imports are absent -- what is checked is that every name the snippets
use exists on our side and lands on the module attribute the reference's trainers read."""
import os
import re
import sys
import types

import pytest

import util

REF = "/root/reference"
DOC = os.path.join(util.ROOT, "INTEGRATION.md")


def _blocks():
    return re.findall(r"```python\n(.*?)```", open(DOC).read(), flags=re.S)


def test_ctypes_stub_of_the_doc_binds_the_built_library():
    from b200unet import _lib
    blocks = [b for b in _blocks() if "C.CDLL(" in b]
    assert blocks, "INTEGRATION.md lost its ctypes stub"
    ns = {}
    exec(blocks[0].replace('"libb2unet.so"', repr(_lib.LIB_PATH)), ns)
    fn = ns["lib"].b2_unet_forward
    assert fn.restype is ns["C"].c_int and list(fn.argtypes) == list(_lib.SIGNATURES["b2_unet_forward"][1])
    # the patch-pipeline stub: structure layout and argument list equal the full binding's
    aug = [b for b in _blocks() if "class AugCase" in b][0]
    head = aug.split("cases = (AugCase")[0]
    ns2 = {"C": ns["C"], "lib": ns["lib"]}
    exec(head, ns2)
    assert [f[0] for f in ns2["AugCase"]._fields_] == [f[0] for f in _lib.AugCase._fields_]
    assert ns["C"].sizeof(ns2["AugCase"]) == ns["C"].sizeof(_lib.AugCase) == 56
    assert len(ns["lib"].b2_aug_crop.argtypes) == len(_lib.SIGNATURES["b2_aug_crop"][1])


@pytest.mark.skipif(not os.path.isdir(REF), reason="the reference is only mounted in the build container")
def test_rebinding_snippets_run_against_the_reference_modules():
    for p in (os.path.join(util.ROOT, "oracle", "shim"), REF):
        if p not in sys.path:
            sys.path.insert(0, p)
    # trainer modules whose nnunet imports are absent here: empty stand-ins that the snippets can assign into
    stand_ins = ["nnunet.training.network_training.nnUNetTrainerV2",
                 "nnunet_ext.training.network_training.multihead.nnUNetTrainerMultiHead",
                 "nnunet_ext.training.network_training.nnViTUNetTrainer"]
    made = []
    for name in stand_ins:
        if name not in sys.modules:
            sys.modules[name] = types.ModuleType(name)
            made.append(name)
            parent, _, leaf = name.rpartition(".")
            try:
                setattr(__import__(parent, fromlist=[leaf]), leaf, sys.modules[name])
            except Exception:
                pass
    import nnunet.network_architecture.generic_UNet as up
    import nnunet_ext.network_architecture.generic_ViT_UNet as vu
    import nnunet_ext.training.loss_functions.deep_supervision as ref_ds
    keep = (up.Generic_UNet, vu.Generic_ViT_UNet, {n: getattr(ref_ds, n) for n in dir(ref_ds) if n.startswith("MultipleOutputLoss")})
    try:
        ran = 0
        for b in _blocks():
            if "up.Generic_UNet =" in b or "vu.Generic_ViT_UNet =" in b or "setattr(ref_ds" in b:
                exec(b, {})
                ran += 1
        assert ran == 3
        from b200unet import deep_supervision as b2_ds
        from b200unet.generic_UNet import Generic_UNet
        from b200unet.generic_ViT_UNet import Generic_ViT_UNet
        assert up.Generic_UNet is Generic_UNet and sys.modules[stand_ins[0]].Generic_UNet is Generic_UNet
        assert sys.modules[stand_ins[1]].Generic_UNet is Generic_UNet
        assert vu.Generic_ViT_UNet is Generic_ViT_UNet and sys.modules[stand_ins[2]].Generic_ViT_UNet is Generic_ViT_UNet
        for n in ("EWC", "RW", "LWF", "MiB", "PLOP", "POD"):
            assert getattr(ref_ds, "MultipleOutputLoss" + n) is getattr(b2_ds, "MultipleOutputLoss" + n)
    finally:
        up.Generic_UNet, vu.Generic_ViT_UNet = keep[0], keep[1]
        for n, v in keep[2].items():
            setattr(ref_ds, n, v)
        for name in made:
            sys.modules.pop(name, None)
