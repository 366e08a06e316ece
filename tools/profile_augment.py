"""20 batches of the GPU patch pipeline at the cfg2 geometry (for an ncu launch list)."""
import os, sys
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "lifelong-nnunet_b200")]
from b200unet import augment
from b200unet.configs import CONFIGS
geom = CONFIGS["cfg2"]
rs = np.random.RandomState(0)
cases = []
for i in range(4):
    sh = tuple(int(p * f) for p, f in zip(geom.patch, (1.3, 1.2, 1.25)))
    d = rs.randn(geom.in_channels + 1, *sh).astype(np.float32)
    d[-1] = (rs.rand(*sh) * 3).astype(np.int64)
    cases.append({"key": "c%d" % i, "data": d})
pipe = augment.GPUPatchPipeline(cases, geom.patch, 2, [(1, 1, 1), (2, 2, 2), (4, 4, 4), (8, 8, 8), (16, 16, 16)], seed=1, prefetch=False)
for _ in range(20):
    next(pipe)
torch.cuda.synchronize()
