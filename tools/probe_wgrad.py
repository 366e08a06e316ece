"""GPU probe (development tool): tensor-core wgrad against torch for both MN-major descriptor stride conventions."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "lifelong-nnunet_b200"))
import torch
import torch.nn.functional as F
from b200unet import ops


def ndhwc(t):
    return t.permute(0, 2, 3, 4, 1).contiguous().to(torch.bfloat16).cuda()


SHAPES = [(2, 8, 16, 16, 32, 32, (1, 1, 1)), (1, 4, 16, 24, 64, 32, (1, 1, 1)), (2, 8, 8, 8, 32, 64, (2, 2, 2)),
          (2, 4, 8, 8, 64, 128, (1, 2, 2)), (2, 4, 4, 4, 320, 320, (1, 1, 1)), (2, 6, 10, 12, 128, 64, (1, 1, 1)),
          (2, 4, 8, 8, 256, 128, (1, 1, 1))]
ops.set_option("tc_wgrad", 1)
for mode in (0, 1):
    ops.set_option("wgrad_desc_mode", mode)
    for (N, D, H, W, cin, cout, stride) in SHAPES:
        g = torch.Generator().manual_seed(5)
        x = torch.randn((N, cin, D, H, W), generator=g)
        w = torch.randn((cout, cin, 3, 3, 3), generator=g) * 0.05
        xr = x.bfloat16().float()
        wr = w.clone().requires_grad_()
        out = F.conv3d(xr, wr, None, stride=stride, padding=1)
        dz = torch.randn(out.shape, generator=g)
        out.backward(dz.bfloat16().float())
        try:
            _, dw, db = ops.conv3d_bwd(ndhwc(x), ndhwc(dz), w.cuda(), stride, need_dx=False)
            torch.cuda.synchronize()
            e = float((dw.cpu() - wr.grad).abs().max() / wr.grad.abs().max())
            eb = float((db.cpu() - dz.bfloat16().float().sum((0, 2, 3, 4))).abs().max() / dz.sum((0, 2, 3, 4)).abs().max())
            print("mode %d shape %s: dw rel err %.3e  db rel err %.3e" % (mode, (N, D, H, W, cin, cout, stride), e, eb), flush=True)
        except Exception as ex:  # noqa
            print("mode %d shape %s: EXC %s" % (mode, (N, D, H, W, cin, cout, stride), ex), flush=True)
