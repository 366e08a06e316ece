"""Times one training step of the ViT-U-Net (cfg4 geometry, V1, ViT-base) -- U-Net parts in the CUDA plan, ViT in ATen."""
import argparse
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "lifelong-nnunet_b200")]

from b200unet import _lib, synth                      # noqa: E402
from b200unet.configs import CONFIGS                  # noqa: E402
from b200unet.trainers import nnUNetTrainerMultiHead  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--config", default="cfg4")
ap.add_argument("--steps", type=int, default=5)
ap.add_argument("--precision", default="bf16")
ap.add_argument("--no-vit", action="store_true")
a = ap.parse_args()
geom = CONFIGS[a.config]
tr = nnUNetTrainerMultiHead(geom, precision=a.precision, use_vit=not a.no_vit)
tr.initialize()
data, targets = synth.make_batch(geom, seed=1)
batch = {"data": data.cuda(), "target": [t.cuda() for t in targets]}


def gen():
    while True:
        yield batch


g = gen()
for _ in range(3):
    tr.run_iteration(g, detach=False)
torch.cuda.synchronize()
l0 = _lib.load().b2_launch_count()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(a.steps):
    l = tr.run_iteration(g, detach=False)
e1.record()
torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / a.steps
# ViT alone (forward + backward through ATen) on a skip-shaped bf16 channels-last tensor
vit_ms = None
if not a.no_vit:
    net = tr.network
    sk = torch.randn(geom.batch, geom.patch[0], geom.patch[1], geom.patch[2], 2 * geom.base_features, device="cuda",
                     dtype=torch.bfloat16 if a.precision == "bf16" else torch.float32)[..., geom.base_features:].permute(0, 4, 1, 2, 3).requires_grad_()
    for i in range(4):
        if i == 1:
            torch.cuda.synchronize(); e0.record()
        with torch.autocast('cuda', dtype=torch.bfloat16, enabled=a.precision == "bf16"):
            o = net.ViT(sk)
        o.float().sum().backward()
    e1.record(); torch.cuda.synchronize()
    vit_ms = e0.elapsed_time(e1) / 3
print(json.dumps({"config": a.config, "vit": not a.no_vit, "precision": a.precision, "ms_per_step": ms,
                  "patches_per_s": geom.batch / ms * 1e3, "vit_fwd_bwd_ms": vit_ms, "loss": float(l),
                  "our_launches_per_step": (_lib.load().b2_launch_count() - l0) / a.steps,
                  "max_mem_gb": torch.cuda.max_memory_allocated() / 2**30}))
