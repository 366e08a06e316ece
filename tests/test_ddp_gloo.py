"""CPU, world_size 2, gloo: host-side logic of the data-parallel path -- one all-reduce (mean) of the
gradient arena per step, the EWC penalty gradient added AFTER the all-reduce (not multiplied by the world size), and
identical parameters / Fisher maps on every rank afterwards.  The arithmetic kernels themselves need a GPU; here the
trainer's synchronisation code runs on a small CPU stand-in network."""
import os
import socket
import sys

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

import util  # noqa: F401


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, out):
    sys.path.insert(0, util.PKG)
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from b200unet.trainers import DataParallelGroup, nnUNetTrainerMultiHead
    torch.manual_seed(0)
    net = torch.nn.Sequential(torch.nn.Linear(6, 5), torch.nn.Tanh(), torch.nn.Linear(5, 3))
    tr = nnUNetTrainerMultiHead.__new__(nnUNetTrainerMultiHead)
    tr.network, tr.ddp = net, DataParallelGroup()
    g = torch.Generator().manual_seed(100 + rank)          # rank-seeded "patches"
    x, y = torch.randn(4, 6, generator=g), torch.randn(4, 3, generator=g)
    loss = ((net(x) - y) ** 2).mean()
    loss.backward()
    local = [p.grad.clone() for p in net.parameters()]
    tr._sync_gradients()
    # penalty gradient (identical on every rank) added after the all-reduce
    star = [p.detach() + 0.1 for p in net.parameters()]
    for p, s in zip(net.parameters(), star):
        p.grad.add_(0.4 * (p.detach() - s))
    out.put((rank, [l for l in local], [p.grad.clone() for p in net.parameters()]))
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.timeout(180)
def test_gradient_allreduce_mean_and_penalty_after_allreduce():
    world, port = 2, _free_port()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = sorted([q.get(timeout=120) for _ in range(world)], key=lambda t: t[0])
    for p in procs:
        p.join(60)
        assert p.exitcode == 0
    (_, l0, s0), (_, l1, s1) = res
    torch.manual_seed(0)
    net = torch.nn.Sequential(torch.nn.Linear(6, 5), torch.nn.Tanh(), torch.nn.Linear(5, 3))
    for a0, a1, b0, b1, p in zip(l0, l1, s0, s1, net.parameters()):
        want = 0.5 * (a0 + a1) + 0.4 * (p.detach() - (p.detach() + 0.1))
        assert torch.allclose(b0, want, atol=1e-6)
        assert torch.equal(b0, b1)            # bit-identical on both ranks -> identical Fisher / parameters
