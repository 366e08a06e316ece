"""cProfile of the e2e loop (pinned host batch -> run_iteration -> loss on host) at cfg2: where the host time of one iteration goes."""
import cProfile, os, pstats, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "lifelong-nnunet_b200")]
import torch
from b200unet import synth
from b200unet.configs import CONFIGS
from b200unet.trainers import nnUNetTrainerEWC
geom = CONFIGS["cfg2"]
tr = nnUNetTrainerEWC(geom, precision="bf16")
tr.initialize()
data, targets = synth.make_batch(geom)
d, t = data.pin_memory(), [x.pin_memory() for x in targets]
fisher, params = synth.make_ewc_state(list(tr.network.named_parameters()))
tr.fisher["task_prev"] = {k: v.cuda() for k, v in fisher.items()}
tr.params["task_prev"] = {k: v.cuda() for k, v in params.items()}
tr.loss.update_ewc_params(tr.fisher, tr.params)
tr.loss.update_network_params(tr.network.named_parameters())
def gen():
    while True:
        yield {'data': d, 'target': t}
g = gen()
for _ in range(8):
    tr.run_iteration(g)
torch.cuda.synchronize()
t0 = time.perf_counter()
for _ in range(50):
    tr.run_iteration(g)
t1 = time.perf_counter()
print("wall ms/iter %.3f" % ((t1 - t0) * 1e3 / 50))
pr = cProfile.Profile()
pr.enable()
for _ in range(50):
    tr.run_iteration(g)
pr.disable()
st = pstats.Stats(pr)
st.sort_stats("tottime").print_stats(22)
