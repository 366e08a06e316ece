"""torch.profiler breakdown of the ATen ViT (forward + backward) fed by a plan-shaped channels-last skip tensor."""
import os
import sys

import torch
from torch.profiler import profile, ProfilerActivity

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "lifelong-nnunet_b200")]
from b200unet.configs import CONFIGS                  # noqa: E402
from b200unet.generic_ViT_UNet import Generic_ViT_UNet  # noqa: E402

geom = CONFIGS["cfg4"]
net = Generic_ViT_UNet(geom.in_channels, geom.base_features, geom.num_classes, geom.num_pool, list(geom.patch),
                       pool_op_kernel_sizes=[list(k) for k in geom.pool], conv_kernel_sizes=[[3, 3, 3]] * (geom.num_pool + 1)).cuda()
sk = torch.randn(geom.batch, *geom.patch, 2 * geom.base_features, device="cuda", dtype=torch.bfloat16)[..., geom.base_features:] \
    .permute(0, 4, 1, 2, 3).requires_grad_()


def step():
    with torch.autocast('cuda', dtype=torch.bfloat16):
        o = net.ViT(sk)
    o.float().sum().backward()


for _ in range(2):
    step()
torch.cuda.synchronize()
with profile(activities=[ProfilerActivity.CUDA, ProfilerActivity.CPU]) as prof:
    step()
    torch.cuda.synchronize()
print(prof.key_averages().table(sort_by="cuda_time_total", row_limit=25, max_name_column_width=70))
