"""Shared helpers for the network golden vectors: deterministic parameter fill (by state_dict key,
so no weights need to be stored) and seeded inputs."""
import zlib

import numpy as np
import torch

SUB = 11   # every 11th element of a flattened output is stored


def fill_state(module, dtype=torch.float64):
    """Overwrite every parameter/buffer with values derived from its key name only."""
    sd = module.state_dict()
    new = {}
    for key, t in sd.items():
        g = torch.Generator().manual_seed(zlib.crc32(key.encode()) & 0x7fffffff)
        if not t.is_floating_point():
            new[key] = t.clone()
            continue
        shape = tuple(t.shape)
        leaf = key.rsplit('.', 1)[-1]
        if leaf == 'running_var':
            v = torch.rand(shape, generator=g, dtype=torch.float64) + 0.5
        elif leaf == 'running_mean':
            v = torch.randn(shape, generator=g, dtype=torch.float64) * 0.1
        elif leaf in ('weight_u', 'weight_v'):
            v = torch.randn(shape, generator=g, dtype=torch.float64)
            v = v / v.norm()
        elif t.dim() >= 2:
            fan_in = max(1, t.numel() // t.shape[0])
            v = torch.randn(shape, generator=g, dtype=torch.float64) * (1.0 / fan_in) ** 0.5
        elif leaf == 'weight':          # norm scale
            v = torch.rand(shape, generator=g, dtype=torch.float64) + 0.5
        else:                           # biases
            v = torch.randn(shape, generator=g, dtype=torch.float64) * 0.05
        new[key] = v.to(dtype)
    module.to(dtype)
    module.load_state_dict(new)
    return module


def rand(seed, *shape):
    return torch.from_numpy(np.random.RandomState(seed).rand(*shape))


def randn(seed, *shape):
    return torch.from_numpy(np.random.RandomState(seed).standard_normal(shape))


def sub(t):
    return t.detach().double().cpu().reshape(-1)[::SUB].numpy().copy()


def run_flownet(net, grad_keys=("conv0.0.weight", "predict_flow2.0.bias", "deconv3.0.weight")):
    """train-mode forward+backward of FlowNet(16) on a seeded (2,3,128,128) input."""
    net.train()
    x = rand(100, 2, 3, 128, 128).to(next(net.parameters()))
    outs = net(x)
    loss = sum((o * randn(101 + i, *o.shape).to(o)).sum() for i, o in enumerate(outs))
    net.zero_grad()
    loss.backward()
    res = {"flow%d" % i: sub(o) for i, o in enumerate(outs)}
    params = dict(net.named_parameters())
    for k in grad_keys:
        res["grad/" + k] = sub(params[k].grad)
    res["bn/conv3.1.running_mean"] = sub(net.state_dict()["conv3.1.running_mean"])
    return res


def run_netd(net):
    net.train()
    x = rand(110, 2, 3, 128, 128).to(next(net.parameters())).requires_grad_(True)
    o = net(x)
    (o * randn(111, *o.shape).to(o)).sum().backward()
    return {"out": sub(o), "grad/x": sub(x.grad), "sn_u": sub(net.state_dict()["nets.1.0.weight_u"])}


def run_lightcnn(net):
    net.eval()
    x = rand(120, 2, 1, 128, 128).to(next(net.parameters())).requires_grad_(True)
    _, fc, pool = net(x)
    ((fc * randn(121, *fc.shape).to(fc)).sum() + (pool * randn(122, *pool.shape).to(pool)).sum()).backward()
    return {"fc": sub(fc), "pool": sub(pool), "grad/x": sub(x.grad)}


def netg_inputs(like):
    x = rand(130, 2, 3, 128, 128).to(like)
    flows = []
    for i, s in enumerate((32, 64, 128)):
        lin = torch.linspace(-1, 1, s, dtype=torch.float64)
        gy, gx = torch.meshgrid(lin, lin, indexing="ij")
        base = torch.stack((gx, gy), 0).unsqueeze(0).repeat(2, 1, 1, 1)
        flows.append((base + 0.08 * randn(131 + i, 2, 2, s, s)).clamp(-1.2, 1.2).to(like).requires_grad_(True))
    return x, flows


def run_netg(net):
    net.train()
    like = next(net.parameters())
    x, flows = netg_inputs(like)
    outs = net(x, flow=flows)
    loss = sum((o * randn(140 + i, *o.shape).to(o)).sum() for i, o in enumerate(outs))
    net.zero_grad()
    loss.backward()
    res = {"rec%d" % i: sub(o) for i, o in enumerate(outs)}
    for i, f in enumerate(flows):
        res["grad/flow%d" % i] = sub(f.grad)
    params = dict(net.named_parameters())
    for k in ("e0.0.weight_orig", "dres2.1.blocks.3.weight_orig", "rec1.0.bias"):
        res["grad/" + k] = sub(params[k].grad)
    res["sn_u/e1.0.weight_u"] = sub(net.state_dict()["e1.0.weight_u"])
    return res


def vgg_torchvision_state(vgg, dtype=torch.float32):
    """The fill-by-key values a torchvision `vgg19().features` state_dict would get, built from the
    shapes of ffwm_b200's VGG19 (keys 'reluX_Y.N.weight' <-> 'features.N.weight')."""
    class _Holder(torch.nn.Module):
        pass
    feats = _Holder()
    feats.features = torch.nn.Module()
    state = {}
    for k, t in vgg.state_dict().items():
        tv_key = "features." + k.split(".", 1)[1]
        g = torch.Generator().manual_seed(zlib.crc32(tv_key.encode()) & 0x7fffffff)
        if t.dim() >= 2:
            fan_in = max(1, t.numel() // t.shape[0])
            v = torch.randn(tuple(t.shape), generator=g, dtype=torch.float64) * (1.0 / fan_in) ** 0.5
        else:
            v = torch.randn(tuple(t.shape), generator=g, dtype=torch.float64) * 0.05
        state[tv_key] = v.to(dtype)
    return state
