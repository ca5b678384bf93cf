"""Spectral normalisation of a whole network in a few batched operations.

The reference wraps every generator / discriminator convolution in `torch.nn.utils.spectral_norm`
(models/base_networks.py:211-216, 374-392; `sn=True`, models/ffwm_model.py:27): each wrapped layer runs, before every
training forward, one power iteration (two mat-vecs, two normalisations, two clones), a dot product and a division —
about a dozen tiny kernels forward and as many backward, 52 layers in netG and 9 in netD (called three times a
step): roughly 1 400 launches per train step for a few MB of arithmetic.

`batch_spectral_norm(net)` keeps the layers exactly as `spectral_norm` made them — same parameters and buffers
(`weight_orig`, `weight_u`, `weight_v`), same state_dict keys, same per-step semantics (one power iteration per
training forward of the network, `u`/`v` updated in place, gradient flowing through sigma) — but takes over their
forward pre-hooks: one pre-hook on the network stacks the weight matrices of all layers of equal shape and runs the
power iteration, sigma and the division once per SHAPE (`torch.bmm` over the stack; netG has 19 distinct shapes,
netD 3).  Arithmetic differs from the per-layer form only in summation order (bmm vs mv).

Default on since round 2 (FFWM_BATCHED_SN=0 in ffwm_b200/base_networks.py restores torch's per-layer hooks).

On a GPU (fp32 weights, one power iteration — the reference's configuration) the whole update runs in the hand-written
kernels of csrc/spectral_norm.cu instead: one device-resident table describes ALL spectrally normalised layers of the
network, and three launches forward + two backward (plus one `cat` of the gradients) replace the ~28 launches per
weight shape of the batched torch formulation — ~780 launches per train step.  FFWM_FUSED_SN=0 keeps the torch path.
"""
import ctypes
import math
import os
import sys

import torch
import torch.nn.functional as F
from torch.autograd.function import once_differentiable

from . import _lib as L

import torch.nn.utils.spectral_norm  # noqa: F401  (makes sure the submodule is loaded)

_SN_CLASS = sys.modules["torch.nn.utils.spectral_norm"].SpectralNorm
FUSED_SN = os.environ.get("FFWM_FUSED_SN", "1") == "1"


def _p(t):
    return ctypes.c_void_p(t.data_ptr())


class _Plan:
    """Device table + sizes for one set of layers (see include/ffwm_b200.h: ffwm_spectral_norm_forward)."""

    def __init__(self, mods, eps):
        self.mods, self.eps = mods, float(eps)
        self.device = mods[0].weight_orig.device
        rows, pre1, pre2, pre3 = [], [0], [0], [0]
        self.slices = []                                   # (element offset, numel, weight shape)
        oe = ot = os_ = 0
        for m in mods:
            w = m.weight_orig
            h, wd = w.shape[0], w[0].numel()
            rows += [w.data_ptr(), m.weight_u.data_ptr(), m.weight_v.data_ptr(), h, wd, oe, ot, os_]
            pre1.append(pre1[-1] + math.ceil(wd / 32))
            pre2.append(pre2[-1] + math.ceil(h / 8))
            pre3.append(pre3[-1] + math.ceil(h * wd / 4096))
            self.slices.append((oe, h * wd, tuple(w.shape)))
            oe, ot, os_ = oe + h * wd, ot + wd, os_ + h
        self.total, self.sum_w, self.sum_h = oe, ot, os_
        self.blocks = (pre1[-1], pre2[-1], pre3[-1])
        self.key = self.pointers(mods)
        self.table = torch.tensor(rows + pre1 + pre2 + pre3, dtype=torch.int64).to(self.device)

    @staticmethod
    def pointers(mods):
        return tuple((m.weight_orig.data_ptr(), m.weight_u.data_ptr(), m.weight_v.data_ptr()) for m in mods)

    def forward(self, update):
        dev, f32 = self.device, torch.float32
        out = torch.empty(self.total, dtype=f32, device=dev)
        t = torch.empty(self.sum_w, dtype=f32, device=dev)
        s = torch.empty(self.sum_h, dtype=f32, device=dev)
        u_saved, v_saved = torch.empty_like(s), torch.empty_like(t)
        sigma = torch.empty(len(self.mods), dtype=f32, device=dev)
        L.call("ffwm_spectral_norm_forward", dev, _p(self.table), len(self.mods), int(bool(update)), ctypes.c_float(self.eps), _p(out), _p(t),
               _p(s), _p(u_saved), _p(v_saved), _p(sigma), *self.blocks)
        return out, u_saved, v_saved, sigma

    def backward(self, grad_flat, u_saved, v_saved, sigma):
        gw = torch.empty_like(grad_flat)
        partial = torch.empty(self.blocks[2], dtype=torch.float64, device=self.device)
        L.call("ffwm_spectral_norm_backward", self.device, _p(self.table), len(self.mods), _p(grad_flat), _p(u_saved), _p(v_saved), _p(sigma),
               _p(gw), _p(partial), self.blocks[2])
        return gw


class _SpectralNormAll(torch.autograd.Function):
    """(weight_orig of every layer) -> (weight_orig / sigma of every layer); u, v are updated in place when `update`."""

    @staticmethod
    def forward(ctx, plan, update, *weights):
        out, u_saved, v_saved, sigma = plan.forward(update)
        ctx.plan = plan
        ctx.save_for_backward(u_saved, v_saved, sigma)
        return tuple(out[o:o + n].view(shape) for o, n, shape in plan.slices)

    @staticmethod
    @once_differentiable
    def backward(ctx, *grads):
        plan = ctx.plan
        parts = [(g.reshape(-1) if g is not None else torch.zeros(n, dtype=torch.float32, device=plan.device))
                 for g, (_, n, _) in zip(grads, plan.slices)]
        gw = plan.backward(torch.cat(parts), *ctx.saved_tensors)
        return (None, None) + tuple(gw[o:o + n].view(shape) for o, n, shape in plan.slices)


class BatchedSpectralNorm:
    def __init__(self, net):
        self.groups = {}
        for m in net.modules():
            for key, hook in list(m._forward_pre_hooks.items()):
                # only the plain case the reference uses: weight matrix = weight.reshape(out, -1), one iteration or more
                if isinstance(hook, _SN_CLASS) and hook.name == "weight" and hook.dim == 0:
                    del m._forward_pre_hooks[key]
                    w = m.weight_orig
                    gkey = (w.shape[0], w[0].numel(), hook.n_power_iterations, hook.eps)
                    self.groups.setdefault(gkey, []).append(m)
        self.n_layers = sum(len(g) for g in self.groups.values())
        self._plan = None
        if self.n_layers:
            net.register_forward_pre_hook(self._update)

    def _fused_ok(self):
        keys = list(self.groups)
        if not (FUSED_SN and keys and all(k[2] == 1 for k in keys) and len({k[3] for k in keys}) == 1):
            return False
        dev = None
        for mods in self.groups.values():
            for m in mods:
                for t in (m.weight_orig, m.weight_u, m.weight_v):
                    if not (t.is_cuda and t.dtype == torch.float32 and t.is_contiguous()) or (dev is not None and t.device != dev):
                        return False
                    dev = t.device
        return True

    def _update_fused(self, net):
        mods = [m for g in self.groups.values() for m in g]
        if self._plan is None or self._plan.key != _Plan.pointers(mods):
            self._plan = _Plan(mods, next(iter(self.groups))[3])
        outs = _SpectralNormAll.apply(self._plan, net.training, *[m.weight_orig for m in mods])
        for m, w in zip(mods, outs):
            setattr(m, "weight", w)

    def _update(self, net, inputs):
        if self._fused_ok():
            return self._update_fused(net)
        for (h, w, n_iter, eps), mods in self.groups.items():
            W = torch.stack([m.weight_orig.reshape(h, w) for m in mods])                  # (L, h, w), differentiable
            U = torch.stack([m.weight_u for m in mods])                                   # (L, h)
            V = torch.stack([m.weight_v for m in mods])                                   # (L, w)
            if net.training:
                with torch.no_grad():
                    Wd = W.detach()
                    for _ in range(n_iter):
                        V = F.normalize(torch.bmm(Wd.transpose(1, 2), U.unsqueeze(-1)).squeeze(-1), dim=1, eps=eps)
                        U = F.normalize(torch.bmm(Wd, V.unsqueeze(-1)).squeeze(-1), dim=1, eps=eps)
                    if n_iter > 0:
                        torch._foreach_copy_([m.weight_u for m in mods], list(U.unbind(0)))
                        torch._foreach_copy_([m.weight_v for m in mods], list(V.unbind(0)))
            sigma = (U * torch.bmm(W, V.unsqueeze(-1)).squeeze(-1)).sum(1)                # u . (W v), per layer
            Wn = W / sigma.view(-1, 1, 1)
            for l, m in enumerate(mods):
                setattr(m, "weight", Wn[l].view_as(m.weight_orig))


def batch_spectral_norm(net):
    """Take over the spectral-norm pre-hooks of `net`'s layers (see the module docstring); returns the manager."""
    return BatchedSpectralNorm(net)
