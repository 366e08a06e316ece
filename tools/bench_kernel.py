"""Run the dominant conv kernel(s) in isolation (for ncu captures and quick CUDA-event timings).
usage: python tools/bench_kernel.py [cin cout D H W B [iters]]"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "lifelong-nnunet_b200"))
import torch
from b200unet import ops

a = [int(v) for v in sys.argv[1:]]
cin, cout, D, H, W, B = (a + [32, 32, 64, 128, 128, 2][len(a):])[:6]
iters = a[6] if len(a) > 6 else 10
dev = torch.device("cuda")
x = torch.randn((B, D, H, W, cin), device=dev).bfloat16()
w = torch.randn((cout, cin, 3, 3, 3), device=dev) * 0.05
b = torch.zeros(cout, device=dev)
dz = torch.randn((B, D, H, W, cout), device=dev).bfloat16()
flops = 2.0 * B * D * H * W * cin * cout * 27


def timeit(fn, n):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n


ms = timeit(lambda: ops.conv3d_fwd(x, w, b, with_stats=False), iters)
print("fwd   %d->%d @ %dx%dx%d B%d: %.3f ms  %.1f TFLOP/s" % (cin, cout, D, H, W, B, ms, flops / ms / 1e9))
ms = timeit(lambda: ops.conv3d_bwd(x, dz, w, need_dx=True), iters)
print("bwd (dgrad+wgrad) : %.3f ms  %.1f TFLOP/s" % (ms, 2 * flops / ms / 1e9))
