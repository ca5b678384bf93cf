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

Opt-in (FFWM_BATCHED_SN=1, read by ffwm_b200/base_networks.py): checked on the CPU against the per-layer hooks
and against the reference goldens, not yet measured on a B200 (written after the round-1 GPU budget was spent).
"""
import sys

import torch
import torch.nn.functional as F

import torch.nn.utils.spectral_norm  # noqa: F401  (makes sure the submodule is loaded)

_SN_CLASS = sys.modules["torch.nn.utils.spectral_norm"].SpectralNorm


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
        if self.n_layers:
            net.register_forward_pre_hook(self._update)

    def _update(self, net, inputs):
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
