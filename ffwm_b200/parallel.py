"""Data-parallel plumbing for the train step: one process per GPU, replicated
weights, per-rank BatchNorm, NCCL all-reduce of the gradients over NVLink/NVSwitch.

The reference has no multi-GPU path; this is new work, kept deliberately small:

* `Distributed`     — thin handle on a `torch.distributed` process group (NCCL on GPUs, gloo in
                      the CPU tests): rank/world, broadcast of module states from rank 0.
* `GradAverager`    — averages the gradients of a fixed parameter list.  Gradients are packed into
                      a few large flat buckets (default 64 MiB: the NVSwitch fabric is not
                      per-link bound, so buckets are sized for launch latency, not link count),
                      each bucket is all-reduced asynchronously while the next one is being
                      packed, and the mean is written back.  Parameters whose `.grad` is None
                      (FlowNet's never-used `inter_conv_occ*`, 6.99 M parameters per net) are
                      skipped — every rank runs the same graph, so the skip pattern is identical
                      on all ranks (asserted once through a checksum of the pattern).
"""
import torch
import torch.distributed as dist


class Distributed:
    def __init__(self, group=None):
        if not dist.is_initialized():
            raise RuntimeError("torch.distributed is not initialised")
        self.group = group
        self.rank = dist.get_rank(group)
        self.world = dist.get_world_size(group)

    def broadcast_module_states(self, modules, src=0):
        """Make parameters and buffers (BN statistics, spectral-norm u/v) identical on all ranks."""
        tensors = []
        for m in modules:
            tensors += [p.data for p in m.parameters()] + [b.data for b in m.buffers() if b.numel() > 0]
        by_dtype = {}
        for t in tensors:
            by_dtype.setdefault((t.dtype, t.device), []).append(t)
        for group in by_dtype.values():
            flat = torch.cat([t.reshape(-1) for t in group])
            dist.broadcast(flat, src=src, group=self.group)
            off = 0
            for t in group:
                t.copy_(flat[off:off + t.numel()].view_as(t))
                off += t.numel()

    def all_reduce_mean_(self, t):
        dist.all_reduce(t, op=dist.ReduceOp.SUM, group=self.group)
        return t.div_(self.world)

    def barrier(self):
        dist.barrier(group=self.group)


class GradAverager:
    def __init__(self, params, distributed, bucket_bytes=64 << 20):
        self.params = [p for p in params]
        self.d = distributed
        self.bucket_bytes = bucket_bytes
        self._plan = None          # list of buckets; bucket = (flat buffer, [(param index, offset, numel)])
        self._pattern = None

    def _make_plan(self, live):
        buckets, cur, cur_bytes = [], [], 0
        for i in live:
            p = self.params[i]
            nbytes = p.numel() * p.element_size()
            if cur and (cur_bytes + nbytes > self.bucket_bytes or self.params[cur[0]].dtype != p.dtype):
                buckets.append(cur)
                cur, cur_bytes = [], 0
            cur.append(i)
            cur_bytes += nbytes
        if cur:
            buckets.append(cur)
        plan = []
        for idxs in buckets:
            p0 = self.params[idxs[0]]
            total = sum(self.params[i].numel() for i in idxs)
            flat = torch.empty(total, dtype=p0.dtype, device=p0.device)
            slots, off = [], 0
            for i in idxs:
                n = self.params[i].numel()
                slots.append((i, off, n))
                off += n
            plan.append((flat, slots))
        return plan

    def average(self):
        """All-reduce (mean) every live gradient in place.  Call after backward, before step."""
        if self.d.world == 1:
            return
        live = tuple(i for i, p in enumerate(self.params) if p.grad is not None)
        if live != self._pattern:
            # identical on every rank?  (sum of a hash must equal world * own hash)
            h = torch.tensor([float(hash(live) % 1000003)], dtype=torch.float64, device=self.params[0].device)
            tot = h.clone()
            dist.all_reduce(tot, group=self.d.group)
            if abs(float(tot) - float(h) * self.d.world) > 0.5:
                raise RuntimeError("GradAverager: ranks disagree on which parameters received gradients")
            self._pattern = live
            self._plan = self._make_plan(live)
        works = []
        inv = 1.0 / self.d.world
        for flat, slots in self._plan:
            views = [flat[off:off + n].view_as(self.params[i]) for i, off, n in slots]
            torch._foreach_copy_(views, [self.params[i].grad for i, _, _ in slots])
            works.append((dist.all_reduce(flat, op=dist.ReduceOp.SUM, group=self.d.group, async_op=True), flat, views, slots))
        for work, flat, views, slots in works:
            work.wait()
            flat.mul_(inv)
            torch._foreach_copy_([self.params[i].grad for i, _, _ in slots], views)

    def live_bytes(self):
        return 0 if self._plan is None else sum(f.numel() * f.element_size() for f, _ in self._plan)
