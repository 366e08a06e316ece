"""GPU debugging aid: layer-by-layer comparison (y, dL/dy) of the CUDA path against the oracle on a small geometry."""
import ctypes as C
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "lifelong-nnunet_b200"))
sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch
from util import cuda_net, oracle_net
from b200unet import _lib, synth
from b200unet.configs import CONFIGS
from oracle import cl_losses

name = sys.argv[1] if len(sys.argv) > 1 else "tiny"
geom = CONFIGS[name]
data, targets = synth.make_batch(geom)
onet = oracle_net(geom)
g = torch.Generator().manual_seed(3)
with torch.no_grad():
    for n, p in onet.named_parameters():
        if "instnorm.weight" in n:
            p.copy_(1 + 0.1 * torch.randn(p.shape, generator=g))
        if "instnorm.bias" in n or "conv.bias" in n:
            p.copy_(0.1 * torch.randn(p.shape, generator=g))
ys, dys, order = {}, {}, []
for n, m in onet.named_modules():
    if n.endswith("lrelu"):
        def hook(mod, inp, out, n=n):
            ys[n] = out.detach().clone()
            order.append(n)
            out.register_hook(lambda gr, n=n: dys.__setitem__(n, gr.detach().clone()))
        m.register_forward_hook(hook)
weights = cl_losses.ds_loss_weights(geom.num_pool)
out = onet(data)
cl_losses.multiple_output_loss2(out, targets, weights).backward()

cnet = cuda_net(geom, onet.state_dict())
co = cnet(data.cuda())
cl_losses.multiple_output_loss2(co, [t.cuda() for t in targets], weights).backward()
torch.cuda.synchronize()
plan = cnet._last_plan
lib = _lib.load()


def view(block, which):
    v = _lib.ActView()
    _lib.check(lib.b2_unet_debug_view(plan.handle, C.c_void_p(plan.workspace.data_ptr()), block, which, C.byref(v)))
    off = (v.ptr - plan.workspace.data_ptr()) // 4
    flat = plan.workspace.view(torch.float32)
    return flat.as_strided((v.n, v.c, v.d, v.h, v.w), (v.d * v.h * v.w * v.pitch, 1, v.h * v.w * v.pitch, v.w * v.pitch, v.pitch), off)


def rel(a, b):
    return float((a.cpu().double() - b.double()).abs().max() / max(float(b.abs().max()), 1e-12))


for i, n in enumerate(order):
    ey = rel(view(i, 1), ys[n])
    edy = rel(view(i, 2), dys[n]) if n in dys else float("nan")
    d = (view(i, 2).cpu() - dys[n]).abs() if n in dys else None
    where = ""
    if d is not None and edy > 1e-4:
        idx = torch.nonzero(d > 0.5 * d.max())
        where = " worst at %s (count>half-max %d) shape %s" % (idx[0].tolist(), idx.shape[0], list(d.shape))
    print("%-50s y %.2e   dy %.2e%s" % (n, ey, edy, where), flush=True)
