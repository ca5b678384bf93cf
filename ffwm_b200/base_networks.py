"""Host-side mirror of the reference's `models/base_networks.py`.

Same public names, constructor / forward signatures and — the checkpoint
contract — the same `state_dict()` keys as the reference, so the released
`latest_net_netG.pth` / `latest_net_flowNetF.pth` load unchanged:

    FlowNet          models/base_networks.py:59-165
    WarpNet          models/base_networks.py:168-173   -> hand-written sm_100a grid_warp kernel
    ResidualBlock / ConvBlock / DeConvBlock / PixelSuffleBlock   :208-272
    FFWM             models/base_networks.py:274-347
    MSDiscriminator  models/base_networks.py:354-437

The networks are described by small tables and assembled by helpers instead of
being spelled out layer by layer.  The feature warp — the op that actually
moves encoder features into the decoder — is
`external_function.grid_warp`, which reads the (B,2,H,W) flow directly: the
reference's `flow.transpose(1,2).transpose(2,3)` copy does not exist here.
Dense convolutions go through `ffwm_b200.conv` (see DESIGN.md for which shapes
run on the hand-written tensor-core path and which stay on cuDNN).
"""
import math
import os

import torch
import torch.nn as nn
import torch.nn.functional as F
from torch.nn.utils import spectral_norm

from . import external_function as EF
from .conv import Conv2d, ConvTranspose2d
from .norm import BatchNorm2d, as_product_norm, fuse_activations
from .spectral import batch_spectral_norm

LRELU_SLOPE = 0.2
# spectral norm of a whole network in one batched update per weight shape (ffwm_b200/spectral.py) instead of one
# power iteration per layer: reference goldens green on CPU and B200, train step 95.2 -> 90.1 ms
# (profiles/r02a_switches.txt).  FFWM_BATCHED_SN=0 restores torch's per-layer hooks for A/B runs.
BATCHED_SN = os.environ.get("FFWM_BATCHED_SN", "1") == "1"


def initialize_msra(modules):
    """Kaiming-normal weights, zero bias for every (transposed) conv (:8-24)."""
    for m in modules:
        if isinstance(m, (nn.Conv2d, nn.ConvTranspose2d)):
            nn.init.kaiming_normal_(m.weight)
            if m.bias is not None:
                nn.init.constant_(m.bias, 0)


# ---------------------------------------------------------------------------------------------
# FlowNet (:30-165).  Every trunk unit is Sequential(conv | deconv, norm, LeakyReLU(0.2)) so the
# keys are `<name>.0.weight`, `<name>.1.running_mean`, ...
# ---------------------------------------------------------------------------------------------
def _unit(cin, cout, norm, k=3, stride=1, transposed=False, bias=True):
    if transposed:
        op = ConvTranspose2d(cin, cout, kernel_size=4, stride=2, padding=1, bias=True)
    else:
        op = Conv2d(cin, cout, kernel_size=k, stride=stride, padding=(k - 1) // 2, bias=bias)
    return nn.Sequential(*fuse_activations([op, as_product_norm(norm)(cout), nn.LeakyReLU(LRELU_SLOPE, inplace=True)]))


def conv(in_planes, out_planes, norm_layer=nn.BatchNorm2d, kernel_size=3, stride=1):
    return _unit(in_planes, out_planes, norm_layer, kernel_size, stride)


def deconv(in_planes, out_planes, norm_layer):
    return _unit(in_planes, out_planes, norm_layer, transposed=True)


def i_conv(in_planes, out_planes, norm_layer, kernel_size=3, stride=1, bias=True):
    return _unit(in_planes, out_planes, norm_layer, kernel_size, stride, bias=bias)


def predict_flow(in_planes):
    """3x3 conv to 2 channels + tanh: an ABSOLUTE sampling grid in [-1,1]."""
    return nn.Sequential(Conv2d(in_planes, 2, kernel_size=3, stride=1, padding=1, bias=True), nn.Tanh())


class FlowNet(nn.Module):
    """FlowNetS-style encoder/decoder; returns (flow128, flow64, flow32) for a 128x128 input."""

    def __init__(self, ngf, norm=nn.BatchNorm2d, x=3):
        super().__init__()
        self.batchNorm = norm
        n = ngf
        # encoder: (name, cin, cout, stride), registration order = reference order
        for name, cin, cout, stride in (
                ("conv0", x, n, 1), ("conv1", n, n, 2), ("conv1_1", n, 2 * n, 1),
                ("conv2", 2 * n, 2 * n, 2), ("conv2_1", 2 * n, 2 * n, 1),
                ("conv3", 2 * n, 4 * n, 2), ("conv3_1", 4 * n, 4 * n, 1),
                ("conv4", 4 * n, 8 * n, 2), ("conv4_1", 8 * n, 8 * n, 1),
                ("conv5", 8 * n, 8 * n, 2), ("conv5_1", 8 * n, 8 * n, 1),
                ("conv6", 8 * n, 16 * n, 2), ("conv6_1", 16 * n, 16 * n, 1)):
            setattr(self, name, conv(cin, cout, norm, stride=stride))
        # decoder level L consumes concat_{L+1}; widths of the concatenations:
        cat_w = {5: 16 * n + 2, 4: 8 * n + 4 * n + 2, 3: 4 * n + 2 * n + 2, 2: n + 2, 1: n // 2 + 2, 0: n // 4 + 2}
        dec_in = {5: 16 * n, 4: cat_w[5], 3: cat_w[4], 2: cat_w[3], 1: cat_w[2], 0: cat_w[1]}
        dec_out = {5: 8 * n, 4: 4 * n, 3: 2 * n, 2: n, 1: n // 2, 0: n // 4}
        for lvl in (5, 4, 3, 2, 1, 0):
            setattr(self, "deconv%d" % lvl, deconv(dec_in[lvl], dec_out[lvl], norm))
        for lvl in (5, 4, 3, 2, 1, 0):
            setattr(self, "inter_conv%d" % lvl, i_conv(cat_w[lvl], dec_out[lvl], norm))
        # occlusion branch of the original network: parameters exist in the checkpoints but the
        # forward pass never uses them — kept for state_dict compatibility
        for lvl in (5, 4, 3, 2, 1, 0):
            setattr(self, "inter_conv_occ%d" % lvl, i_conv(cat_w[lvl] - 1, dec_out[lvl], norm))
        head_in = {6: 16 * n, 5: 8 * n, 4: 4 * n, 3: 2 * n, 2: n, 1: n // 2, 0: n // 4}
        for lvl in (6, 5, 4, 3, 2, 1, 0):
            setattr(self, "predict_flow%d" % lvl, predict_flow(head_in[lvl]))
        for lvl in (6, 5, 4, 3, 2, 1):
            setattr(self, "upsampled_flow%d_to_%d" % (lvl, lvl - 1), ConvTranspose2d(2, 2, 4, 2, 1))
        initialize_msra(self.modules())

    def forward(self, x):
        enc = {0: self.conv0(x)}
        for lvl in range(1, 7):
            down = getattr(self, "conv%d" % lvl)(enc[lvl - 1])
            enc[lvl] = getattr(self, "conv%d_1" % lvl)(down)

        flow = self.predict_flow6(enc[6])
        feat = enc[6]                       # what the next deconv consumes
        flows = {}
        for lvl in (5, 4, 3, 2, 1, 0):
            up_flow = getattr(self, "upsampled_flow%d_to_%d" % (lvl + 1, lvl))(flow)
            up_feat = getattr(self, "deconv%d" % lvl)(feat)
            parts = (enc[lvl], up_feat, up_flow) if lvl >= 3 else (up_feat, up_flow)   # skips only down to 16x16
            feat = torch.cat(parts, 1)
            flow = getattr(self, "predict_flow%d" % lvl)(getattr(self, "inter_conv%d" % lvl)(feat))
            flows[lvl] = flow
        return flows[0], flows[1], flows[2]


class WarpNet(nn.Module):
    """`F.grid_sample(images, flow as (B,H,W,2), bilinear, zeros, align_corners=False)` (:168-173)
    on the hand-written kernel; `mode='nearest'` is not used anywhere in the reference."""

    def forward(self, images, flow, mode='bilinear'):
        if mode != 'bilinear':
            raise NotImplementedError("WarpNet: only mode='bilinear' is used by FFWM")
        return EF.grid_warp(images.contiguous(), flow.contiguous())


# ---------------------------------------------------------------------------------------------
# Generator building blocks (:179-272)
# ---------------------------------------------------------------------------------------------
class Tanh2(nn.Module):
    def __init__(self):
        super().__init__()
        self.tanh = nn.Tanh()

    def forward(self, x):
        return (self.tanh(x) + 1) / 2


_ACTIVATIONS = {'relu': nn.ReLU, 'lrelu': lambda: nn.LeakyReLU(LRELU_SLOPE), 'sigmoid': nn.Sigmoid,
                'tanh': nn.Tanh, 'tanh2': Tanh2}
_NORMS = {'bn': BatchNorm2d, 'in': nn.InstanceNorm2d}


def get_activ(name):
    if name not in _ACTIVATIONS:
        raise NotImplementedError('Activation %s not implemented' % name)
    return _ACTIVATIONS[name]()


def get_norm(name, ch):
    if name not in _NORMS:
        raise NotImplementedError('Normalization %s not implemented' % name)
    return _NORMS[name](ch)


def _maybe_sn(layer, sn):
    return spectral_norm(layer) if sn else layer


class ResidualBlock(nn.Module):
    """activ(blocks(x) + input(x)); `input` is a 1x1 projection (:208-233).  Without spectral norm
    the reference pads by `kernel` (not kernel//2), growing the map — reproduced as is."""

    def __init__(self, inc, outc=None, kernel=3, stride=1, activ='lrelu', norm='bn', sn=False):
        super().__init__()
        outc = inc // stride if outc is None else outc
        pad = kernel // 2 if sn else kernel
        self.activ = get_activ(activ)
        self.input = _maybe_sn(Conv2d(inc, outc, 1, 1, padding=0), sn)
        self.blocks = nn.Sequential(*fuse_activations([_maybe_sn(Conv2d(inc, outc, kernel, 1, pad), sn), get_norm(norm, outc),
                                                       nn.LeakyReLU(LRELU_SLOPE),
                                                       _maybe_sn(Conv2d(outc, outc, kernel, 1, pad), sn), get_norm(norm, outc)]))

    def forward(self, x):
        last = self.blocks[-1]
        if type(self.activ) is nn.LeakyReLU and isinstance(last, BatchNorm2d) and last.act_slope is None:
            # the closing batch norm adds the projection and applies the activation in its own kernels (ffwm_b200/norm.py)
            h = x
            for m in list(self.blocks)[:-1]:
                h = m(h)
            return last(h, residual=self.input(x), act_slope=self.activ.negative_slope)
        return self.activ(self.blocks(x) + self.input(x))


def _block(first, outc, activ, norm, res, resk, bn, sn):
    seq = list(first)
    if bn:
        seq.append(get_norm(norm, outc))
    if activ is not None:
        seq.append(get_activ(activ))
    seq = fuse_activations(seq)
    seq += [ResidualBlock(outc, activ=activ, kernel=resk, norm=norm, sn=sn) for _ in range(res)]
    return nn.Sequential(*seq)


def ConvBlock(inc, outc, ks=3, s=1, p=0, activ='lrelu', norm='bn', res=0, resk=3, bn=True, sn=False):
    return _block([_maybe_sn(Conv2d(inc, outc, ks, s, p), sn)], outc, activ, norm, res, resk, bn, sn)


def DeConvBlock(inc, outc, ks=3, s=1, p=0, op=0, activ='relu', norm='bn', res=0, resk=3, bn=True, sn=False):
    return _block([_maybe_sn(ConvTranspose2d(inc, outc, ks, s, p, op), sn)], outc, activ, norm, res, resk, bn, sn)


def PixelSuffleBlock(inc, outc, ks=3, s=1, p=0, activ='lrelu', norm='bn', res=0, bn=True, sn=False):
    """3x3 conv to 4*outc channels + PixelShuffle(2); ks/s/p are accepted and ignored, as in the reference."""
    return _block([_maybe_sn(Conv2d(inc, outc * 4, 3, 1, 1), sn), nn.PixelShuffle(2)], outc, activ, norm, res, 3, bn, sn)


class FFWM(nn.Module):
    """Generator (:274-347): 4-level encoder; per decoder level: pixel-shuffle upsample, WARP the
    encoder feature with the flow of that scale, flip-concat, attention gate, concat the
    upsampled low-res reconstruction, two residual blocks, 3x3 sigmoid reconstruction."""

    def __init__(self, num_layers=3, isflip=True, sn=False):
        super().__init__()
        enc_ch = [64, 64, 128, 256]
        dec_ch = [256, 128, 64, 64]
        self.isflip = isflip
        dm = 3 if isflip else 2          # decoder width multiplier: warped (+ flipped) + decoded
        am = dm - 1                      # attention width multiplier
        self.layers = num_layers

        self.e0 = ConvBlock(3, enc_ch[0], 7, 1, 3, res=1, bn=False, sn=sn)
        for i in (1, 2, 3):
            setattr(self, "e%d" % i, ConvBlock(enc_ch[i - 1], enc_ch[i], 4, 2, 1, res=1, sn=sn))
        # widths entering the residual stacks: level 0 has no low-res reconstruction to concatenate
        res_w = [dec_ch[1] * dm, dec_ch[2] * dm + 3, dec_ch[3] * dm + 3]
        d_in = [dec_ch[0], res_w[0], res_w[1]]
        for i in range(3):
            setattr(self, "d%d" % i, PixelSuffleBlock(d_in[i], dec_ch[i + 1], 4, 2, 1, sn=sn))
        for i in range(3):
            setattr(self, "dres%d" % i, nn.Sequential(*[ResidualBlock(res_w[i], activ='lrelu', sn=sn) for _ in range(2)]))
        for i in range(3):
            setattr(self, "rec%d" % i, ConvBlock(res_w[i], 3, 3, 1, 1, bn=False, activ='sigmoid', sn=sn))
        for i in range(3):
            w = enc_ch[2 - i] * am
            setattr(self, "att%d" % i, nn.Sequential(ConvBlock(w, w, 3, 1, 1, sn=sn),
                                                     ResidualBlock(w, w, activ='sigmoid', sn=sn)))
        self.warpNet = WarpNet()
        if sn and BATCHED_SN:
            batch_spectral_norm(self)

    def forward(self, x, flow=None, return_att=False):
        fencs = [self.e0(x)]
        for i in range(1, self.layers + 1):
            fencs.append(getattr(self, "e%d" % i)(fencs[-1]))
        fdec, recons, att = fencs[-1], [], None
        for i in range(self.layers):
            dec = getattr(self, "d%d" % i)(fdec)
            warped = self.warpNet(fencs[self.layers - 1 - i], flow[i])
            skip = torch.cat((warped, torch.flip(warped, (3,))), 1) if self.isflip else warped
            att = getattr(self, "att%d" % i)(skip)
            parts = [skip * att, dec]
            if recons:   # TP-GAN style: feed the previous scale's reconstruction, upsampled x2
                parts.append(F.interpolate(recons[-1], scale_factor=2, mode='bilinear'))
            fdec = getattr(self, "dres%d" % i)(torch.cat(parts, 1))
            recons.append(getattr(self, "rec%d" % i)(fdec))
        if return_att:
            return recons[-3], recons[-2], recons[-1], att
        return recons[-3], recons[-2], recons[-1]


# ---------------------------------------------------------------------------------------------
# Multi-scale discriminator (:354-437), InGAN style
# ---------------------------------------------------------------------------------------------
class MSDiscriminator(nn.Module):
    def __init__(self, real_crop_size, inc=3, max_n_scales=9, scale_factor=2, base_channels=64,
                 extra_conv_layers=0, sigmoid=True):
        super().__init__()
        self.inc, self.base_channels, self.scale_factor = inc, base_channels, scale_factor
        self.min_size = 16
        self.extra_conv_layers = extra_conv_layers
        self.sigmoid = sigmoid
        smallest = min(real_crop_size) if hasattr(real_crop_size, "__len__") else real_crop_size
        # as many scales as fit the real examples: ceil(log_sf(size / min_size))
        self.max_n_scales = min(int(math.ceil(math.log(smallest * 1.0 / self.min_size) / math.log(scale_factor))),
                                max_n_scales)
        self.nets = nn.ModuleList([self.make_net() for _ in range(self.max_n_scales)])
        if BATCHED_SN:
            batch_spectral_norm(self)

    def make_net(self):
        c = self.base_channels
        layers = []
        for cin, cout in ((self.inc, c), (c, 2 * c), (2 * c, 4 * c)):      # three stride-2 SN conv blocks
            layers += [spectral_norm(Conv2d(cin, cout, kernel_size=3, stride=2, padding=1)),
                       BatchNorm2d(cout), nn.LeakyReLU(LRELU_SLOPE, True)]
        for _ in range(self.extra_conv_layers):
            layers += [spectral_norm(Conv2d(2 * c, 2 * c, kernel_size=3, bias=True)),
                       BatchNorm2d(2 * c), nn.LeakyReLU(LRELU_SLOPE, True)]
        if self.sigmoid:
            layers += [spectral_norm(Conv2d(4 * c, 1, kernel_size=1)), nn.Sigmoid()]
        else:
            layers += [Conv2d(4 * c, 1, kernel_size=1)]
        return nn.Sequential(*fuse_activations(layers))

    def forward(self, input_tensor):
        total = self.nets[0](input_tensor)
        size = total.shape[2:]
        # (the reference's scale_weights has five entries, so at most five scales are ever summed)
        for i, net in enumerate(list(self.nets[1:5]), start=1):
            small = F.interpolate(input_tensor, scale_factor=self.scale_factor ** (-i), mode='bilinear')
            total = total + F.interpolate(net(small), size=size, mode='bilinear')
        return total
