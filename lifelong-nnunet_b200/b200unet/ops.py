"""Thin Python wrappers over the building-block entry points of include/b2unet.h (used by the per-kernel parity tests
and by bench.py's dominant-kernel timing).  Tensors are NDHWC ("channels last 3d"), fp32 or bf16."""
import ctypes as C

import torch

from . import _lib


def _stream(dev):
    return C.c_void_p(torch.cuda.current_stream(dev).cuda_stream)


def _desc(x, cout, stride, out_pitch=None):
    d = _lib.ConvDesc()
    d.n, d.d, d.h, d.w, d.cin = (int(s) for s in x.shape)
    d.cout = int(cout)
    for i in range(3):
        d.stride[i] = int(stride[i])
    d.in_pitch = int(x.stride(3))
    d.out_pitch = int(out_pitch or cout)
    d.dtype = _lib.B2_F32 if x.dtype == torch.float32 else _lib.B2_BF16
    return d


def set_option(name, value):
    _lib.check(_lib.load().b2_set_option(name.encode(), int(value)))


def conv3d_fwd(x, w_pt, bias, stride=(1, 1, 1), eps=1e-5, with_stats=True):
    """x: (N,D,H,W,Cin) NDHWC; w_pt: (Cout,Cin,3,3,3) fp32; returns z (N,Do,Ho,Wo,Cout) and stats (N,Cout,2)."""
    lib = _lib.load()
    N, D, H, W, _ = x.shape
    cout = w_pt.shape[0]
    od, oh, ow = [(s - 1) // st + 1 for s, st in zip((D, H, W), stride)]
    z = torch.empty((N, od, oh, ow, cout), dtype=x.dtype, device=x.device)
    stats = torch.empty((N, cout, 2), dtype=torch.float32, device=x.device) if with_stats else None
    d = _desc(x, cout, stride)
    scr = torch.empty(int(lib.b2_conv3d_scratch_bytes(C.byref(d))), dtype=torch.uint8, device=x.device)
    _lib.check(lib.b2_conv3d_fwd(C.byref(d), x.data_ptr(), w_pt.data_ptr(), None if bias is None else bias.data_ptr(),
                                 z.data_ptr(), None if stats is None else stats.data_ptr(), float(eps), scr.data_ptr(),
                                 _stream(x.device)))
    return z, stats


def conv3d_bwd(x, dz, w_pt, stride=(1, 1, 1), need_dx=True, accumulate_into=None):
    lib = _lib.load()
    cout, cin = w_pt.shape[0], w_pt.shape[1]
    d = _desc(x, cout, stride)
    dx = None
    if need_dx:
        dx = accumulate_into if accumulate_into is not None else torch.empty_like(x)
    dw = torch.empty_like(w_pt)
    db = torch.empty(cout, dtype=torch.float32, device=x.device)
    scr = torch.empty(int(lib.b2_conv3d_scratch_bytes(C.byref(d))), dtype=torch.uint8, device=x.device)
    _lib.check(lib.b2_conv3d_bwd(C.byref(d), x.data_ptr(), dz.data_ptr(), w_pt.data_ptr(),
                                 None if dx is None else dx.data_ptr(), int(accumulate_into is not None), dw.data_ptr(),
                                 db.data_ptr(), scr.data_ptr(), _stream(x.device)))
    return dx, dw, db
