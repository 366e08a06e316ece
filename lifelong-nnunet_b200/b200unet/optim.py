"""Fused optimiser / importance-update entry points (multi-tensor CUDA kernels behind include/b2unet.h).

* ``B2SGD.clip_and_step(max_norm)`` replaces ``clip_grad_norm_(params, 12)`` + ``optimizer.step()`` of the reference
  iteration (reference nnUNetTrainerMultiHead.py:629-630/640-641) for the optimiser built at :294-301
  (SGD lr 1e-2, wd 3e-5, momentum 0.99, nesterov).
* ``fisher_square`` replaces ewc:298-304 (F = grad^2), ``rw_update`` replaces rw:240-262.
"""
import ctypes as C

import torch

from . import _lib


def _stream(dev):
    return C.c_void_p(torch.cuda.current_stream(dev).cuda_stream)


class B2SGD(torch.optim.Optimizer):
    def __init__(self, params, lr, weight_decay=0.0, momentum=0.0, nesterov=False):
        super().__init__(params, dict(lr=lr, weight_decay=weight_decay, momentum=momentum, nesterov=nesterov))
        self.last_grad_norm = None

    @torch.no_grad()
    def clip_and_step(self, max_norm=12.0):
        lib = _lib.load()
        for group in self.param_groups:
            ps = [p for p in group['params'] if p.grad is not None]
            if not ps:
                continue
            dev = ps[0].device
            table = (_lib.SgdEntry * len(ps))()
            total = 0
            for i, p in enumerate(ps):
                st = self.state[p]
                if 'momentum_buffer' not in st:
                    # torch.optim.SGD creates the buffer lazily PER PARAMETER as clone(d_p); a zero buffer run through
                    # the warm update gives momentum * 0 + d_p = d_p, bit for bit, so parameters that first receive a
                    # gradient in a later task (zero-weight head, ViT-bypassed bottleneck) need no special launch
                    st['momentum_buffer'] = torch.zeros_like(p)
                if not p.grad.is_contiguous():
                    p.grad = p.grad.contiguous()
                table[i].theta, table[i].grad = p.data_ptr(), p.grad.data_ptr()
                table[i].momentum, table[i].numel = st['momentum_buffer'].data_ptr(), p.numel()
                total += p.numel()
            first = 0
            norm = torch.empty(1, dtype=torch.float32, device=dev)
            scr = torch.empty(int(lib.b2_sgd_scratch_bytes(len(ps), total)), dtype=torch.uint8, device=dev)
            _lib.check(lib.b2_sgd_clip_step(table, len(ps), float(group['lr']), float(group['momentum']),
                                            float(group['weight_decay']), int(group['nesterov']), float(max_norm),
                                            first, norm.data_ptr(), scr.data_ptr(), _stream(dev)))
            self.last_grad_norm = norm

    def step(self, closure=None):
        if closure is not None:
            raise NotImplementedError
        self.clip_and_step(float('inf'))


@torch.no_grad()
def fisher_square(grads):
    """F_i = g_i^2 for a list of gradient tensors, one launch (ewc:303)."""
    lib = _lib.load()
    dev = grads[0].device
    out = [torch.empty_like(g) for g in grads]
    n = len(grads)
    gp = (C.c_void_p * n)(*[g.data_ptr() for g in grads])
    fp = (C.c_void_p * n)(*[f.data_ptr() for f in out])
    ne = (C.c_int64 * n)(*[g.numel() for g in grads])
    scr = torch.empty(int(lib.b2_multitensor_scratch_bytes(n)), dtype=torch.uint8, device=dev)
    _lib.check(lib.b2_fisher_square(gp, fp, ne, n, scr.data_ptr(), _stream(dev)))
    return out


@torch.no_grad()
def rw_update(params, prev, fisher, score, alpha, eps=1e-8, have_prev=True):
    """In-place Riemannian-walk update of (score, prev, fisher) for params with gradients (rw:240-262)."""
    lib = _lib.load()
    dev = params[0].device
    n = len(params)
    table = (_lib.RwEntry * n)()
    for i, p in enumerate(params):
        table[i].theta, table[i].grad = p.data_ptr(), p.grad.data_ptr()
        table[i].prev, table[i].fisher, table[i].score = prev[i].data_ptr(), fisher[i].data_ptr(), score[i].data_ptr()
        table[i].numel = p.numel()
    scr = torch.empty(int(lib.b2_multitensor_scratch_bytes(n)), dtype=torch.uint8, device=dev)
    _lib.check(lib.b2_rw_update(table, n, float(alpha), float(eps), int(bool(have_prev)), scr.data_ptr(), _stream(dev)))
