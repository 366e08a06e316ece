"""Timing probe of conv_halo_kernel with parts of the pipeline switched off (b2_set_option("halo_dbg", mask)):
1 no global stores, 2 no MMA issue, 4 no TMA slab loads, 8 no TMEM loads.  Results are WRONG by construction -- timing only."""
import ctypes as C
import os
import sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "lifelong-nnunet_b200"))
import torch
from b200unet import _lib, ops
lib = _lib.load()
dev = torch.device("cuda")
a = [int(v) for v in sys.argv[1:]]
cin, cout, D, H, W, B = (a + [32, 32, 64, 128, 128, 2][len(a):])[:6]
x = torch.randn((B, D, H, W, cin), device=dev).bfloat16()
z = torch.empty((B, D, H, W, cout), device=dev, dtype=torch.bfloat16)
w = torch.randn((cout, cin, 3, 3, 3), device=dev) * 0.05
bias = torch.zeros(cout, device=dev)
desc = ops._desc(x, cout, (1, 1, 1))
scr = torch.empty(int(lib.b2_conv3d_scratch_bytes(C.byref(desc))), dtype=torch.uint8, device=dev)
shadow = torch.empty(int(lib.b2_conv3d_shadow_bytes(C.byref(desc))), dtype=torch.uint8, device=dev)
st = C.c_void_p(torch.cuda.current_stream(dev).cuda_stream)
_lib.check(lib.b2_conv3d_make_shadow(C.byref(desc), w.data_ptr(), shadow.data_ptr(), st))
flops = 2.0 * B * D * H * W * cin * cout * 27
def run(stats, n=10):
    fn = lib.b2_conv3d_fwd_shadow_stats if stats else lib.b2_conv3d_fwd_shadow
    for _ in range(3):
        _lib.check(fn(C.byref(desc), x.data_ptr(), shadow.data_ptr(), bias.data_ptr(), z.data_ptr(), scr.data_ptr(), st))
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n):
        _lib.check(fn(C.byref(desc), x.data_ptr(), shadow.data_ptr(), bias.data_ptr(), z.data_ptr(), scr.data_ptr(), st))
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n * 1e3
for mask in [int(m) for m in os.environ.get('MASKS', '0,1,2,4,8,3,6,5,9,12,14,7,15').split(',')]:
    ops.set_option("halo_dbg", mask)
    print("%d->%d dbg %3d [%s%s%s%s]: no-stats %7.1f us   stats %7.1f us" % (cin, cout, mask, "S" if mask & 1 else "-", "M" if mask & 2 else "-",
          "T" if mask & 4 else "-", "L" if mask & 8 else "-", run(False), run(True)), flush=True)
ops.set_option("halo_dbg", 0)
