"""Data-parallel plumbing on CPU: world_size 2, gloo."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from ffwm_b200.parallel import Distributed, GradAverager
        torch.manual_seed(100 + rank)                     # different init per rank on purpose
        net = torch.nn.Sequential(torch.nn.Linear(5, 7), torch.nn.BatchNorm1d(7), torch.nn.Linear(7, 3))
        unused = torch.nn.Linear(4, 4)                    # never receives a gradient (like inter_conv_occ*)
        d = Distributed()
        d.broadcast_module_states([net, unused])
        w0 = net[0].weight.detach().clone()
        params = list(net.parameters()) + list(unused.parameters())
        avg = GradAverager(params, d, bucket_bytes=64)    # tiny buckets: several all-reduces in flight
        results = []
        for step in range(2):
            for p in params:
                p.grad = None
            x = torch.full((4, 5), float(rank + 1 + step))
            net(x).sum().backward()
            local = [p.grad.numpy().copy() for p in net.parameters()]
            avg.average()
            results.append((local, [p.grad.numpy().copy() for p in net.parameters()]))
        assert all(p.grad is None for p in unused.parameters())
        q.put((rank, w0.numpy(), results))     # numpy: plain pickling, no shared-memory handles
    finally:
        dist.destroy_process_group()


def test_grad_averager_and_broadcast_world2():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    out = sorted([q.get(timeout=120) for _ in procs], key=lambda t: t[0])
    for p in procs:
        p.join(60)
        assert p.exitcode == 0
    (_, w_a, res_a), (_, w_b, res_b) = out
    assert (w_a == w_b).all()                          # broadcast made the replicas identical
    for step in range(2):
        (loc_a, avg_a), (loc_b, avg_b) = res_a[step], res_b[step]
        for la, lb, ga, gb in zip(loc_a, loc_b, avg_a, avg_b):
            assert abs(ga - (la + lb) / 2).max() <= 1e-6 * max(1.0, abs(ga).max())
            assert (ga == gb).all()


def test_single_process_is_a_noop():
    class One:
        world, rank, group = 1, 0, None
    from ffwm_b200.parallel import GradAverager
    p = torch.nn.Parameter(torch.ones(3))
    p.grad = torch.full((3,), 2.0)
    GradAverager([p], One()).average()
    assert torch.equal(p.grad, torch.full((3,), 2.0))
