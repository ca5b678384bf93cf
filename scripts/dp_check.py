"""2+ GPU check of the data-parallel train step (run under torchrun): replicas stay identical, and the overlapped
averager (all-reduces captured inside the step's CUDA graph) computes what the pack-after-backward averager computes."""
import os
import sys

import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from benchmarks.train_step import make_batch  # noqa: E402


def run(overlap, graph, rank, dev):
    os.environ["FFWM_DP_OVERLAP"] = "1" if overlap else "0"
    from ffwm_b200.parallel import Distributed
    from ffwm_b200.train_step import FFWMTrainer
    torch.manual_seed(0)
    tr = FFWMTrainer(dev, distributed=Distributed(), graph=graph)
    batches = [{k: (v.to(dev) if torch.is_tensor(v) else v) for k, v in make_batch(2, 50 + 10 * i + rank).items()} for i in range(4)]
    if graph:
        tr.enable_cuda_graph(batches[0], warmup=2)
    else:
        for _ in range(2):
            tr.step(batches[0])
    for b in batches[1:]:
        tr.step(b)
    torch.cuda.synchronize()
    sums = torch.tensor([float(sum(p.detach().double().abs().sum() for p in getattr(tr, n).parameters()))
                         for n in ("netG", "netD", "flowNetF", "flowNetB")], dtype=torch.float64, device=dev)
    gathered = [torch.zeros_like(sums) for _ in range(dist.get_world_size())]
    dist.all_gather(gathered, sums)
    assert all(torch.equal(g, gathered[0]) for g in gathered), "replicas diverged (overlap=%s graph=%s)" % (overlap, graph)
    return sums.cpu(), tr.get_current_losses()


def main():
    rank, local = int(os.environ["RANK"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist.init_process_group("nccl", device_id=dev)
    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cuda.matmul.allow_tf32 = False
    res = {}
    for overlap, graph in ((False, False), (True, False), (True, True)):
        res[(overlap, graph)] = run(overlap, graph, rank, dev)
        if rank == 0:
            print("overlap=%s graph=%s |param| sums %s loss_G %.5f" % (overlap, graph, res[(overlap, graph)][0].tolist(), res[(overlap, graph)][1]["loss_G"]), flush=True)
    base = res[(False, False)][0]
    for k, (s, _) in res.items():
        rel = float(((s - base).abs() / base).max())
        assert rel <= 1e-4, (k, rel)          # three Adam steps amplify summation-order noise (overlap vs not: other bucket order)
    if rank == 0:
        print("dp_check: ok (world %d)" % dist.get_world_size())
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
