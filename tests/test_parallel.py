"""Data-parallel plumbing on CPU: world_size 2, gloo."""
import copy
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


def _worker(rank, world, port, q, overlap):
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
        avg = GradAverager(params, d, bucket_bytes=64, overlap=overlap)    # tiny buckets: several all-reduces in flight
        results = []
        for step in range(3):
            avg.zero()                                    # (overlap: zero-fills the persistent buckets from step 2 on)
            x = torch.full((4, 5), float(rank + 1 + step)) + torch.arange(20.).view(4, 5) * (rank + 1)
            # this rank's own gradient, from a hook-free twin (in overlap mode the buckets are already being reduced
            # while backward runs, so p.grad cannot be read "before the average")
            twin = copy.deepcopy(net)
            local = [g.numpy().copy() for g in torch.autograd.grad(twin(x).pow(2).sum(), list(twin.parameters()))]
            net(x).pow(2).sum().backward()
            avg.average()
            results.append((local, [p.grad.numpy().copy() for p in net.parameters()]))
        assert all(p.grad is None for p in unused.parameters())
        if overlap:                                       # the gradients live in the flat buckets: no pack / unpack copies
            flats = [f for f, _ in avg._plan]
            assert all(any(p.grad.data_ptr() >= f.data_ptr() and p.grad.data_ptr() < f.data_ptr() + f.numel() * 4 for f in flats)
                       for p in net.parameters())
        q.put((rank, w0.numpy(), results))     # numpy: plain pickling, no shared-memory handles
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("overlap", [True, False], ids=["overlapped-buckets", "pack-after-backward"])
def test_grad_averager_and_broadcast_world2(overlap):
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q, overlap)) for r in range(2)]
    for p in procs:
        p.start()
    out = sorted([q.get(timeout=120) for _ in procs], key=lambda t: t[0])
    for p in procs:
        p.join(60)
        assert p.exitcode == 0
    (_, w_a, res_a), (_, w_b, res_b) = out
    assert (w_a == w_b).all()                          # broadcast made the replicas identical
    for step in range(3):
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


def _trainer_worker(rank, world, port, q):
    """One data-parallel FFWM train step per rank on the CPU (the warps served by torch ops through the oracle's
    cpu_ops shim — test infrastructure), different data and different initial weights per rank."""
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    torch.set_num_threads(max(1, (os.cpu_count() or 2) // world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from oracle import train_cpu
        from ffwm_b200.parallel import Distributed
        from ffwm_b200.train_step import FFWMTrainer
        torch.manual_seed(100 + rank)                     # replicas start different: the broadcast must fix that
        with train_cpu.cpu_ops():
            tr = FFWMTrainer("cpu", distributed=Distributed())
            g = torch.Generator().manual_seed(7 + rank)   # per-rank data
            batch = {'img_S': torch.rand(1, 3, 128, 128, generator=g), 'img_F': torch.rand(1, 3, 128, 128, generator=g),
                     'mask_F': (torch.rand(1, 1, 128, 128, generator=g) > 0.3).float(),
                     'mask_S': (torch.rand(1, 1, 128, 128, generator=g) > 0.3).float(),
                     'lm_F': torch.randint(20, 108, (1, 1000, 2), generator=g), 'titers': 30000, 'epoch': 0}
            tr.set_input(batch)
            tr.optimize_parameters()
        sums = {}
        for name in ("netG", "netD", "flowNetF", "flowNetB"):
            net = getattr(tr, name)
            sums[name] = float(sum(p.detach().double().sum() for p in net.parameters()))
            sums[name + "_abs"] = float(sum(p.detach().double().abs().sum() for p in net.parameters()))
            sums[name + "_grad"] = float(sum(p.grad.double().abs().sum() for p in net.parameters() if p.grad is not None))
        unused = [n for n, p in tr.flowNetF.named_parameters() if p.grad is None]
        q.put((rank, sums, float(tr.loss_G), float(tr.loss_D), unused))
    finally:
        dist.destroy_process_group()


def test_data_parallel_train_step_world2_keeps_replicas_identical():
    """SURVEY 8e on the CPU: after one step on different per-rank batches the two replicas hold bit-identical
    weights and gradients (broadcast at start, averaged gradients, same Adam update), although their losses differ."""
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_trainer_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    out = sorted([q.get(timeout=600) for _ in procs], key=lambda t: t[0])
    for p in procs:
        p.join(120)
        assert p.exitcode == 0
    (_, s0, lg0, ld0, un0), (_, s1, lg1, ld1, un1) = out
    assert s0 == s1, "replicas diverged: %r vs %r" % (s0, s1)
    assert lg0 != lg1 and ld0 != ld1                      # the ranks really saw different data
    assert all(v > 0 for k, v in s0.items() if k.endswith("_grad"))
    assert un0 == un1 and any("inter_conv_occ" in n for n in un0)    # never-used parameters are skipped on both ranks
