"""Host-side enqueue time of one training step (is the step CPU-bound?): per-step host timestamps of run_iteration with no
synchronisation -- the first steps after a sync show the pure enqueue cost before the launch queue fills."""
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "lifelong-nnunet_b200"))
import torch
from b200unet import synth
from b200unet.configs import CONFIGS
from b200unet.trainers import nnUNetTrainerEWC

geom = CONFIGS["cfg2"]
tr = nnUNetTrainerEWC(geom, precision="bf16")
tr.initialize()
data, targets = synth.make_batch(geom)
d, t = data.cuda(), [x.cuda() for x in targets]
named = list(tr.network.named_parameters())
fisher, params = synth.make_ewc_state(named)
tr.fisher["task_prev"] = {k: v.cuda() for k, v in fisher.items()}
tr.params["task_prev"] = {k: v.cuda() for k, v in params.items()}
tr.loss.update_ewc_params(tr.fisher, tr.params)
tr.loss.update_network_params(tr.network.named_parameters())


def gen():
    while True:
        yield {'data': d, 'target': t}


g = gen()
for _ in range(5):
    tr.run_iteration(g, detach=False)
torch.cuda.synchronize()
for rep in range(2):
    ts = [time.perf_counter()]
    for _ in range(12):
        tr.run_iteration(g, detach=False)
        ts.append(time.perf_counter())
    torch.cuda.synchronize()
    te = time.perf_counter()
    print("host ms per step:", " ".join("%.2f" % ((b - a) * 1e3) for a, b in zip(ts, ts[1:])), "| total incl. drain %.2f ms" % ((te - ts[0]) * 1e3))
