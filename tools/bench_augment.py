"""Throughput of the GPU patch pipeline (b200unet/augment.py) at a BASELINE geometry: patches/s of the pipeline alone (device
timed, CUDA events) and of pipeline -> trainer.run_iteration (EWC, bf16).  Synthetic cases resident in HBM.
usage: python tools/bench_augment.py [cfg2] [batches]"""
import json
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "lifelong-nnunet_b200")]
from b200unet import augment, synth  # noqa: E402
from b200unet.configs import CONFIGS  # noqa: E402
from b200unet.trainers import nnUNetTrainerEWC  # noqa: E402

name = sys.argv[1] if len(sys.argv) > 1 else "cfg2"
n = int(sys.argv[2]) if len(sys.argv) > 2 else 50
geom = CONFIGS[name]
rs = np.random.RandomState(0)
cases = []
for i in range(6):                                      # hippocampus-like cases somewhat larger than the patch
    sh = tuple(int(p * f) for p, f in zip(geom.patch, (1.3, 1.2, 1.25)))
    d = rs.randn(geom.in_channels + 1, *sh).astype(np.float32)
    seg = (rs.rand(*[s // 8 + 1 for s in sh]) * geom.num_classes).astype(np.int64).repeat(8, 0).repeat(8, 1).repeat(8, 2)[:sh[0], :sh[1], :sh[2]]
    d[-1] = seg
    cases.append({"key": "c%d" % i, "data": d})
strides, cum = [(1, 1, 1)], [1, 1, 1]
for k in geom.pool[:-1]:
    cum = [a * b for a, b in zip(cum, k)]
    strides.append(tuple(cum))
pipe = augment.GPUPatchPipeline(cases, geom.patch, geom.batch, strides, seed=1)
for _ in range(5):
    next(pipe)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
launches = 0
e0.record()
for _ in range(n):
    next(pipe)
    launches += pipe.launches_last
e1.record()
torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / n
res = {"workload": name, "patch": list(geom.patch), "generator_patch": list(pipe.gen_patch), "batch": geom.batch,
       "pipeline_ms_per_batch": ms, "pipeline_patches_per_s": geom.batch * 1000.0 / ms, "launches_per_batch": launches / n}
tr = nnUNetTrainerEWC(geom, precision="bf16", task="B")
tr.initialize()
fisher, params = synth.make_ewc_state(list(tr.network.named_parameters()))
tr.fisher["A"] = {k: v.to(tr.device) for k, v in fisher.items()}
tr.params["A"] = {k: v.to(tr.device) for k, v in params.items()}
tr.loss.update_ewc_params(tr.fisher, tr.params)
for _ in range(5):
    tr.run_iteration(pipe)
torch.cuda.synchronize()
e0.record()
for _ in range(n):
    tr.run_iteration(pipe)
e1.record()
torch.cuda.synchronize()
ms2 = e0.elapsed_time(e1) / n
res.update({"train_with_pipeline_ms_per_step": ms2, "train_with_pipeline_patches_per_s": geom.batch * 1000.0 / ms2})
print(json.dumps(res))
