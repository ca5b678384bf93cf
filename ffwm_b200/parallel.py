"""Data-parallel plumbing for the train step: one process per GPU, replicated
weights, per-rank BatchNorm, NCCL all-reduce of the gradients over NVLink/NVSwitch.

The reference has no multi-GPU path; this is new work, kept deliberately small:

* `Distributed`     — thin handle on a `torch.distributed` process group (NCCL on GPUs, gloo in
                      the CPU tests): rank/world, broadcast of module states from rank 0.
* `GradAverager`    — averages the gradients of a fixed parameter list.  Parameters whose `.grad` is None
                      (FlowNet's never-used `inter_conv_occ*`, 6.99 M parameters per net) are
                      skipped — every rank runs the same graph, so the skip pattern is identical
                      on all ranks (asserted once through a checksum of the pattern).
                      overlap=True (default): after the first backward the live gradients LIVE in a few
                      persistent flat buckets (`p.grad` is a view; no pack / unpack copies, NCCL's AVG instead
                      of a scaling pass), ordered by the order in which autograd finished them.
                        in_backward=True (eager steps): a post-accumulate hook launches a bucket's all-reduce
                          the moment its last gradient is written — while the rest of the backward pass is
                          still running; `average()` then only waits for the collectives in flight.
                        in_backward=False (CUDA-graph steps): the backward pass is a captured graph, so the
                          buckets (one per dtype by default: a single 428 MB all-reduce for netG + both
                          FlowNets) are reduced by `average()` between two graph replays.  Capturing the
                          collectives INTO the graph was measured on 2 x B200 (round 2): no faster than this
                          (67.3 vs 66.6 ms) and the replicas diverged, so it is not done.
                      overlap=False: gradients are packed into flat buckets after the backward pass, each
                      bucket is all-reduced asynchronously while the next one is being packed, and the mean is
                      copied back (round 1's path; fully exposed, kept as the fallback FFWM_DP_OVERLAP=0).
                      Buckets default to 32 MiB: the NVSwitch fabric is not per-link bound, so they are sized
                      for launch latency and for how much of the last bucket stays exposed, not for link count.
"""
import os

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
    def __init__(self, params, distributed, bucket_bytes=None, overlap=None, in_backward=True):
        self.params = [p for p in params]
        self.d = distributed
        self.in_backward = bool(in_backward)
        self.bucket_bytes = bucket_bytes if bucket_bytes is not None else ((32 << 20) if in_backward else (1 << 40))
        self.overlap = (os.environ.get("FFWM_DP_OVERLAP", "1") == "1") if overlap is None else bool(overlap)
        self._plan = None          # list of buckets; bucket = (flat buffer, [(param index, offset, numel)])
        self._pattern = None
        self._order = []           # overlap: parameter indices in the order autograd finished them (first backward)
        self._bucket_of = {}       # overlap: param index -> bucket index
        self._pending = []         # overlap: gradients still missing per bucket in the current backward
        self._works = []
        if self.overlap and self.d.world > 1:
            for i, p in enumerate(self.params):
                if p.requires_grad:
                    p.register_post_accumulate_grad_hook(self._make_hook(i))

    # ------------------------------------------------------------------ shared
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

    def _check_pattern(self, live):
        # identical on every rank?  (sum of a hash must equal world * own hash)
        h = torch.tensor([float(hash(live) % 1000003)], dtype=torch.float64, device=self.params[0].device)
        tot = h.clone()
        dist.all_reduce(tot, group=self.d.group)
        if abs(float(tot) - float(h) * self.d.world) > 0.5:
            raise RuntimeError("GradAverager: ranks disagree on which parameters received gradients")

    def _all_reduce_mean(self, flat):
        """Asynchronous mean all-reduce of one bucket; returns a callable that completes it on the current stream."""
        if flat.is_cuda:
            work = dist.all_reduce(flat, op=dist.ReduceOp.AVG, group=self.d.group, async_op=True)
            return work.wait
        work = dist.all_reduce(flat, op=dist.ReduceOp.SUM, group=self.d.group, async_op=True)     # gloo has no AVG
        inv = 1.0 / self.d.world

        def done():
            work.wait()
            flat.mul_(inv)
        return done

    # ------------------------------------------------------------------ overlap mode
    def _make_hook(self, i):
        def hook(_param):
            if self._plan is None:
                self._order.append(i)                  # first backward: learn the order, reduce afterwards
                return
            b = self._bucket_of.get(i)
            if b is None:
                raise RuntimeError("GradAverager: a parameter outside the recorded pattern received a gradient")
            self._pending[b] -= 1
            if self._pending[b] == 0 and self.in_backward:
                self._works.append(self._all_reduce_mean(self._plan[b][0]))
        return hook

    def zero(self):
        """Replaces optimizer.zero_grad() for these parameters: zero-fills the persistent buckets (gradients stay views
        into them) once they exist, drops the gradients before that."""
        if not (self.overlap and self._plan is not None):
            for p in self.params:
                p.grad = None
            self._order = []
            return
        torch._foreach_zero_([flat for flat, _ in self._plan])
        self._pending = [len(slots) for _, slots in self._plan]
        self._works = []

    def _adopt(self):
        """After the first backward: buckets in completion order, gradients moved into them for good."""
        seen, order = set(), []
        for i in self._order:
            if i not in seen and self.params[i].grad is not None:
                seen.add(i)
                order.append(i)
        order += [i for i, p in enumerate(self.params) if p.grad is not None and i not in seen]
        live = tuple(order)
        self._check_pattern(live)                      # same gradients in the same completion order on every rank
        self._pattern = live
        self._plan = self._make_plan(live)
        for b, (flat, slots) in enumerate(self._plan):
            views = [flat[off:off + n].view_as(self.params[i]) for i, off, n in slots]
            torch._foreach_copy_(views, [self.params[i].grad for i, _, _ in slots])
            for (i, _, _), v in zip(slots, views):
                self.params[i].grad = v
                self._bucket_of[i] = b
        self._pending = [0] * len(self._plan)
        self._works = [self._all_reduce_mean(flat) for flat, _ in self._plan]

    # ------------------------------------------------------------------ the call after backward
    def average(self):
        """All-reduce (mean) every live gradient in place.  Call after backward, before step."""
        if self.d.world == 1:
            return
        if self.overlap:
            if self._plan is None:
                self._adopt()
            elif not self.in_backward:
                self._works = [self._all_reduce_mean(flat) for flat, _ in self._plan]     # (hooks do not run in a graph replay)
            elif any(self._pending):
                raise RuntimeError("GradAverager: %d bucket(s) never completed in this backward pass (call zero() before "
                                   "every backward)" % sum(1 for n in self._pending if n))
            for done in self._works:
                done()
            self._works = []
            return
        live = tuple(i for i, p in enumerate(self.params) if p.grad is not None)
        if live != self._pattern:
            self._check_pattern(live)
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
