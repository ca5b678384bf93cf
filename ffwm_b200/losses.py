"""Host-side mirror of the reference's `models/losses.py` (SURVEY.md 8a a15-a16 and the callers
of the custom ops): same class names, constructor / call signatures and numerics.

    GANLoss                          models/losses.py:7-59
    LandmarkLoss / MultiScaleLDLoss  :61-74, :114-126
    IdentityLoss                     :76-112   (LightCNN features, optional centre crop by grid warp)
    MSL1Loss                         :130-155  (illumination loss: 3 grid warps by the reverse flow)
    (Multi)AffineRegularizationLoss  :163-223  (block_extractor + local_attn_reshape — K4..K7)
    VGG19 / VGGLoss / StyleLoss / PerceptualLoss   :225-319, :398-519
    PerceptualCorrectness            :322-396  (resample2d — K1..K3 — or the bilinear grid warp)

Every warp goes through the hand-written kernels (`external_function`, `base_networks.WarpNet`).
Deliberate, documented differences:
  * integer landmark division is floor division on every torch version (the reference relies on
    torch<=1.5 `LongTensor.div`, which true-divides on torch 2.x and breaks `gather`);
  * VGG19 takes its ImageNet weights from a local torchvision state_dict file when one is given
    (`FFWM_VGG19_WEIGHTS` or the `weights_path` argument) and is otherwise randomly initialised —
    there is no network here; the architecture, layer names and keys are the reference's.
"""
import os

import numpy as np
import torch
import torch.nn as nn
import torch.nn.functional as F

from . import base_networks, ops
from .conv import LOSSNET_MATH_FWD, Conv2d, set_forward_math
from .pool import MaxPool2d
from .external_function import AffineResidualFunction, BlockExtractor, LocalAttnReshape, Resample2d, grid_warp

# the affine regularisation as one kernel per direction (csrc/affine_reg.cu) and the correlation column-max of
# PerceptualCorrectness as a fused tcgen05 GEMM + running max (csrc/corr_max.cu) instead of the reference's chains of
# library ops; FFWM_FUSED_AFFINE=0 / FFWM_FUSED_CORRMAX=0 restore those chains (A/B runs, and what the CPU tests exercise).
FUSED_AFFINE = os.environ.get("FFWM_FUSED_AFFINE", "1") == "1"
FUSED_CORRMAX = os.environ.get("FFWM_FUSED_CORRMAX", "1") == "1"
# PerceptualLoss.many: the L1 terms of all pairs of a resolution in one batched pass per VGG layer (FFWM_BATCHED_L1=0: per pair)
BATCHED_L1 = os.environ.get("FFWM_BATCHED_L1", "1") == "1"


class GANLoss(nn.Module):
    _WITH_TARGET = ('lsgan', 'vanilla', 'nagan')    # ('nagan' is the reference's spelling, :38)

    def __init__(self, gan_mode, target_real_label=1.0, target_fake_label=0.0):
        super().__init__()
        self.register_buffer('real_label', torch.tensor(target_real_label))
        self.register_buffer('fake_label', torch.tensor(target_fake_label))
        self.gan_mode = gan_mode
        if gan_mode == 'nsgan':
            self.criterion = nn.BCELoss()
        elif gan_mode == 'lsgan':
            self.loss = nn.MSELoss()
        elif gan_mode == 'vanilla':
            self.loss = nn.BCEWithLogitsLoss()
        elif gan_mode == 'hinge':
            self.criterion = nn.ReLU()
        elif gan_mode in ('wgangp', 'dcgan'):
            self.loss = None
        else:
            raise NotImplementedError('gan mode %s not implemented' % gan_mode)

    def get_target_tensor(self, prediction, target_is_real):
        return (self.real_label if target_is_real else self.fake_label).expand_as(prediction)

    def __call__(self, predictions, target_is_real, for_dis=None, weights=None):
        if type(predictions) is not list:
            predictions = [predictions]
        loss = 0
        for p in predictions:
            if self.gan_mode in self._WITH_TARGET:
                loss += self.loss(p, self.get_target_tensor(p, target_is_real))
            elif self.gan_mode == 'wgangp':
                loss += -p.mean() if target_is_real else p.mean()
            elif self.gan_mode == 'hinge':
                if for_dis:
                    loss += self.criterion(1 + (-p if target_is_real else p)).mean()
                else:
                    loss += (-p).mean()
            elif self.gan_mode == 'dcgan':
                loss += torch.mean(F.softplus(-p if target_is_real else p))
        return loss


# ------------------------------------------------------------------------------------------
class LandmarkLoss(nn.Module):
    """MSE between the flow sampled at the frontal landmarks and the profile landmarks in [-1,1]."""

    def __init__(self):
        super().__init__()
        self.criterionL2 = nn.MSELoss()

    def forward(self, flow, lm_S, lm_F, gate):
        b, _, s, _ = flow.size()
        flat = flow.permute(0, 2, 3, 1).reshape(b, -1, 2)
        index = (lm_F[:, :, 0:1] + lm_F[:, :, 1:2] * s).expand(-1, -1, 2)
        flow_points = torch.gather(flat, 1, index)
        gt_points = lm_S.float() / (s / 2.0) - 1
        return self.criterionL2(flow_points * gate, gt_points * gate)


def _int_div(t, scale):
    # torch 1.5 semantics of LongTensor.div(int)
    return torch.div(t, scale, rounding_mode='floor') if not t.is_floating_point() else t.div(scale)


class MultiScaleLDLoss(nn.Module):
    def __init__(self):
        super().__init__()
        self.criterionLD = LandmarkLoss()
        self.weights = [1000, 1000, 1500]
        self.img_size = 128

    def forward(self, flows, lm_S, lm_F, gate):
        total = 0
        for w, flow in zip(self.weights, flows):
            scale = self.img_size // flow.size(3)
            total += w * self.criterionLD(flow, _int_div(lm_S, scale), _int_div(lm_F, scale), gate)
        return total


class IdentityLoss(nn.Module):
    def __init__(self, lightcnn, crop=False):
        super().__init__()
        self.lightcnn = set_forward_math(lightcnn, LOSSNET_MATH_FWD)     # frozen feature extractor (ffwm_b200/conv.py)
        self.criterionL1 = nn.L1Loss()
        self.warpNet = base_networks.WarpNet()
        self.crop = crop

    def forward(self, out, gt):
        if self.crop:   # centre 98x98 face crop via a fixed sampling grid, resized back
            grid = self.build_grid(out.size(0), 98).type_as(out)
            size = (out.size(2), out.size(3))
            out = F.interpolate(self.warpNet(out, grid), size, mode='bilinear')
            gt = F.interpolate(self.warpNet(gt, grid), size, mode='bilinear')
        _, fc_out, pool_out = self.lightcnn(out.mean(dim=1, keepdim=True))
        with torch.no_grad():
            _, fc_gt, pool_gt = self.lightcnn(gt.mean(dim=1, keepdim=True))
        return self.criterionL1(fc_out, fc_gt.detach()) + self.criterionL1(pool_out, pool_gt.detach())

    def many(self, pairs):
        """[self(out, gt) for out, gt in pairs] with one LightCNN pass over all the generated images and one (without
        autograd) over all the targets.  LightCNN-29 has no batch statistics: per-sample features, and each pair's
        two L1 means over its own slice, are unchanged."""
        outs, gts = [], []
        for out, gt in pairs:
            if self.crop:
                grid = self.build_grid(out.size(0), 98).type_as(out)
                size = (out.size(2), out.size(3))
                out = F.interpolate(self.warpNet(out, grid), size, mode='bilinear')
                gt = F.interpolate(self.warpNet(gt, grid), size, mode='bilinear')
            outs.append(out.mean(dim=1, keepdim=True))
            gts.append(gt.mean(dim=1, keepdim=True))
        sizes = [o.size(0) for o in outs]
        _, fc_out, pool_out = self.lightcnn(torch.cat(outs))
        with torch.no_grad():
            _, fc_gt, pool_gt = self.lightcnn(torch.cat(gts))
        return [self.criterionL1(a, b) + self.criterionL1(c, d)
                for a, b, c, d in zip(fc_out.split(sizes), fc_gt.split(sizes), pool_out.split(sizes), pool_gt.split(sizes))]

    def build_grid(self, b, d):
        """(b,2,d,d) absolute grid centred on pixel (64,77) of a 128x128 image (:101-111)."""
        r = d // 2
        line = torch.linspace(-r, r, d)
        gx = (line.view(1, d).expand(d, d) + (64 - 64)) / 64
        gy = (line.view(d, 1).expand(d, d) + (77 - 64)) / 64
        return torch.stack((gx, gy), 0).unsqueeze(0).repeat(b, 1, 1, 1)


class MSL1Loss(nn.Module):
    def __init__(self, criterionL1):
        super().__init__()
        self.warpNet = base_networks.WarpNet()
        self.criterionL1 = criterionL1
        self.l1_weights = [1, 1, 1.5]

    def resize_as(self, img, tar, mode='bilinear'):
        size = tar.shape[2:]
        if mode == 'bilinear':
            return F.interpolate(img, size, mode=mode, align_corners=True)
        return F.interpolate(img, size, mode=mode)

    def forward(self, flows, img_Ss, img_F, mask=None):
        total = 0
        for w, flow, img_S in zip(self.l1_weights, flows, img_Ss):
            target = self.resize_as(img_F, flow)
            warped = self.warpNet(img_S, flow)
            if mask is not None:
                m = self.resize_as(mask, flow, 'nearest')
                warped, target = warped * m, target * m
            total += w * self.criterionL1(warped, target)
        return total


# ------------------------------------------------------------------------------------------
# Affine regularisation (GFLA): the only live user of block_extractor / local_attn_reshape
# ------------------------------------------------------------------------------------------
class AffineRegularizationLoss(nn.Module):
    """For every kz x kz window of the sampling grid, the residual of the best affine fit:
    g^T (K^T K) g with K = A (A^T A)^-1 A^T - I  (:163-223)."""

    def __init__(self, kz):
        super().__init__()
        self.kz = kz
        self.criterion = nn.L1Loss()
        self.extractor = BlockExtractor(kernel_size=kz)
        self.reshape = LocalAttnReshape()
        idx = np.arange(kz)
        A = np.ones([kz * kz, 3])
        A[:, 0] = np.repeat(idx, kz)
        A[:, 1] = np.tile(idx, kz)
        K = A @ np.linalg.inv(A.T @ A) @ A.T - np.identity(kz * kz)
        self.kernel = torch.from_numpy(K.T @ K).view(kz * kz, 1, kz, kz)

    def __call__(self, flow_fields):
        grid = self.flow2grid(flow_fields)
        weights = self.kernel.type_as(flow_fields)
        return self.calculate_loss(grid[:, 0:1], weights) + self.calculate_loss(grid[:, 1:2], weights)

    def calculate_loss(self, grid, weights):
        if FUSED_AFFINE and grid.is_cuda and self.kz in (3, 5, 7):
            # the whole chain below is w^T (K^T K) w / kz^2 per window: one kernel per direction (csrc/affine_reg.cu)
            q = weights.reshape(self.kz ** 2, self.kz ** 2).contiguous()
            return torch.mean(AffineResidualFunction.apply(grid, q, self.kz)) * self.kz ** 2
        results = F.conv2d(grid, weights)                       # [b, kz*kz, h', w']
        b, _, h, w = results.size()
        kernels_new = self.reshape(results, self.kz)            # K6: [b,1,kz*h',kz*w']
        f = results.new_full((b, 2, h, w), float(self.kz // 2))
        grid_H = self.extractor(grid, f)                        # K4 as an unfold
        result = F.avg_pool2d(grid_H * kernels_new, self.kz, self.kz)
        return torch.mean(result) * self.kz ** 2

    def flow2grid(self, flow_field):
        return flow_field.add(1.0).div(2.0).mul(128.0)


class MultiAffineRegularizationLoss(nn.Module):
    def __init__(self, kz_dic):
        super().__init__()
        self.kz_dic = kz_dic
        self.method_dic = {key: AffineRegularizationLoss(kz) for key, kz in kz_dic.items()}
        self.layers = sorted(kz_dic, reverse=True)

    def __call__(self, flow_fields):
        return sum(self.method_dic[self.layers[i]](f) for i, f in enumerate(flow_fields))


# ------------------------------------------------------------------------------------------
# VGG19 feature pyramid and the perceptual losses
# ------------------------------------------------------------------------------------------
class VGG19(nn.Module):
    """torchvision VGG19 `features[:36]` cut into the 16 named relu stages; sub-module names are the
    original feature indices (`relu3_1.9`, `relu3_1.10`, ...) exactly as in the reference (:398-519)."""

    CFG = (64, 64, 'M', 128, 128, 'M', 256, 256, 256, 256, 'M', 512, 512, 512, 512, 'M', 512, 512, 512, 512)

    def __init__(self, weights_path=None):
        super().__init__()
        layers, cin = [], 3
        for v in self.CFG:
            if v == 'M':
                layers.append(MaxPool2d(kernel_size=2, stride=2))
            else:
                layers += [Conv2d(cin, v, kernel_size=3, padding=1), nn.ReLU(inplace=True)]
                cin = v
        self.stage_names = []
        block, conv_i, idx = 1, 0, 0
        while idx < len(layers):
            seq = nn.Sequential()
            if isinstance(layers[idx], nn.MaxPool2d):          # a pool opens the next block
                seq.add_module(str(idx), layers[idx])
                idx += 1
                block, conv_i = block + 1, 0
            seq.add_module(str(idx), layers[idx])
            seq.add_module(str(idx + 1), layers[idx + 1])
            idx += 2
            conv_i += 1
            name = 'relu%d_%d' % (block, conv_i)
            self.add_module(name, seq)
            self.stage_names.append(name)
        weights_path = weights_path or os.environ.get('FFWM_VGG19_WEIGHTS')
        if weights_path:
            self.load_torchvision(torch.load(weights_path, map_location='cpu'))
        for p in self.parameters():
            p.requires_grad = False
        set_forward_math(self, LOSSNET_MATH_FWD)                          # frozen feature extractor (ffwm_b200/conv.py)

    def load_torchvision(self, state):
        own = {k.split('.', 1)[1]: k for k in self.state_dict()}       # '9.weight' -> 'relu3_1.9.weight'
        self.load_state_dict({own[k[len('features.'):]]: v for k, v in state.items()
                              if k.startswith('features.') and k[len('features.'):] in own})

    def forward(self, x):
        out = {}
        for name in self.stage_names:
            x = getattr(self, name)(x)
            out[name] = x
        return out


def _gram(x):
    b, ch, h, w = x.size()
    f = x.view(b, ch, w * h)
    return f.bmm(f.transpose(1, 2)) / (h * w * ch)


class VGGLoss(nn.Module):
    def __init__(self, weights=[1.0, 1.0, 1.0, 1.0, 1.0]):
        super().__init__()
        self.add_module('vgg', VGG19())
        self.criterion = nn.L1Loss()
        self.weights = weights

    def compute_gram(self, x):
        return _gram(x)

    def __call__(self, x, y):
        xv, yv = self.vgg(x), self.vgg(y)
        content = sum(w * self.criterion(xv['relu%d_1' % (i + 1)], yv['relu%d_1' % (i + 1)])
                      for i, w in enumerate(self.weights))
        style = sum(self.criterion(_gram(xv[k]), _gram(yv[k])) for k in ('relu2_2', 'relu3_4', 'relu4_4', 'relu5_2'))
        return content, style


class StyleLoss(nn.Module):
    def __init__(self):
        super().__init__()
        self.add_module('vgg', VGG19())
        self.criterion = nn.L1Loss()

    def compute_gram(self, x):
        return _gram(x)

    def __call__(self, x, y):
        xv, yv = self.vgg(x), self.vgg(y)
        return sum(self.criterion(_gram(xv[k]), _gram(yv[k])) for k in ('relu2_2', 'relu3_4', 'relu4_4', 'relu5_2'))


class PerceptualLoss(nn.Module):
    def __init__(self, layers=['relu1_1', 'relu2_1', 'relu3_1', 'relu4_1', 'relu5_1'],
                 weights=[1.0, 1.0 / 2, 1.0 / 4, 1.0 / 4, 1.0 / 8]):
        super().__init__()
        self.add_module('vgg', VGG19())
        self.criterion = nn.L1Loss()
        self.weights = weights
        self.layers = layers

    def __call__(self, x, y):
        xv, yv = self.vgg(x), self.vgg(y)
        loss = 0.0
        for layer, w in zip(self.layers, self.weights):
            loss += w * self.criterion(xv[layer], yv[layer].detach())
        return loss

    def many(self, pairs):
        """[self(x, y) for x, y in pairs] with one VGG pass over all the x of a resolution and one (without autograd:
        the targets are detached anyway) over all the y, instead of two passes per pair.  VGG19 has no batch
        statistics, so every sample's features — and every pair's loss, reduced over that pair's slice exactly as
        above — are unchanged; the train step's 14 passes (10 of them on 32x32 inputs, batch 8) become 6."""
        out = [0.0] * len(pairs)
        by_shape = {}
        for i, (x, _) in enumerate(pairs):
            by_shape.setdefault(tuple(x.shape[1:]), []).append(i)
        for idxs in by_shape.values():
            sizes = [pairs[i][0].size(0) for i in idxs]
            xv = self.vgg(torch.cat([pairs[i][0] for i in idxs]))
            with torch.no_grad():
                yv = self.vgg(torch.cat([pairs[i][1] for i in idxs]))
            if BATCHED_L1 and isinstance(self.criterion, nn.L1Loss) and self.criterion.reduction == "mean":
                # every pair's mean |a - b| per layer from ONE pass over the concatenated features: per-sample means, then a
                # (samples x pairs) averaging matrix — the samples of a group have equal numel, so the mean of a pair's
                # per-sample means IS its L1Loss; 5 layers x 7 pairs x (sub, abs, mean, mul, add) and their autograd
                # mirror become 5 x 4 kernels
                seg = self._segment_matrix(sizes, xv[self.layers[0]])
                acc = None
                for layer, w in zip(self.layers, self.weights):
                    per = F.l1_loss(xv[layer], yv[layer], reduction="none").flatten(1).mean(1)      # (samples,)
                    acc = w * (per @ seg) if acc is None else acc + w * (per @ seg)
                for k, i in enumerate(idxs):
                    out[i] = out[i] + acc[k]
                continue
            for layer, w in zip(self.layers, self.weights):
                for i, a, b in zip(idxs, xv[layer].split(sizes), yv[layer].split(sizes)):
                    out[i] = out[i] + w * self.criterion(a, b)
        return out

    def _segment_matrix(self, sizes, like):
        device = like.device
        key = (tuple(sizes), str(device), like.dtype)
        cache = self.__dict__.setdefault("_seg_cache", {})
        if key not in cache:
            m = torch.zeros(sum(sizes), len(sizes), dtype=like.dtype)
            r = 0
            for k, n in enumerate(sizes):
                m[r:r + n, k] = 1.0 / n
                r += n
            cache[key] = m.to(device)
        return cache[key]


class PerceptualCorrectness(nn.Module):
    """Sampling-correctness loss of GFLA (:322-396): cosine similarity between the warped source
    features and the target features, normalised by the best similarity any source location
    reaches (a column max of the [N^2 x N^2] correlation)."""

    def __init__(self, layer=['relu1_1', 'relu2_1', 'relu3_1', 'relu4_1']):
        super().__init__()
        self.add_module('vgg', VGG19())
        self.layer = layer
        self.eps = 1e-8
        self.resample = Resample2d(4, 1, sigma=2)
        self.l1_loss = nn.L1Loss()

    def __call__(self, target, source, flow_list, used_layers, norm_mask=None, use_bilinear_sampling=True):
        used_layers = sorted(used_layers, reverse=True)
        self.target_vgg, self.source_vgg = self.vgg(target), self.vgg(source)
        loss = 0
        for flow, used in zip(flow_list, used_layers):
            loss += self.calculate_loss(flow, self.layer[used], norm_mask, use_bilinear_sampling)
        return loss

    def _sample(self, source_vgg, flow, use_bilinear_sampling):
        return self.bilinear_warp(source_vgg, flow) if use_bilinear_sampling else self.resample(source_vgg, flow)

    def calculate_loss(self, flow, layer, norm_mask=None, use_bilinear_sampling=False):
        target_vgg, source_vgg = self.target_vgg[layer], self.source_vgg[layer]
        b, c, h, w = target_vgg.shape
        flow = F.interpolate(flow, [h, w])
        target_all = target_vgg.view(b, c, -1)                          # [b, C, N2]
        if (FUSED_CORRMAX and ops.corr_max_supported(target_vgg) and source_vgg.shape == target_vgg.shape
                and not (source_vgg.requires_grad or target_vgg.requires_grad)):
            # normalisation + [N2 x N2] product + max over the source axis without materialising anything (csrc/corr_max.cu)
            correction_max = ops.corr_max(source_vgg, target_vgg, self.eps)
        else:
            source_all = source_vgg.view(b, c, -1).transpose(1, 2)      # [b, N2, C]
            source_norm = source_all / (source_all.norm(dim=2, keepdim=True) + self.eps)
            target_norm = target_all / (target_all.norm(dim=1, keepdim=True) + self.eps)
            correction_max = self._column_max(source_norm, target_norm)     # [b, N2]
        input_sample = self._sample(source_vgg, flow, use_bilinear_sampling).view(b, c, -1)
        correction_sample = F.cosine_similarity(input_sample, target_all)
        loss_map = torch.exp(-correction_sample / (correction_max + self.eps))
        e1 = torch.exp(torch.tensor(-1).type_as(loss_map))
        if norm_mask is None:
            return torch.mean(loss_map) - e1
        norm_mask = F.interpolate(norm_mask, size=(h, w)).view(-1, h * w)
        return (torch.sum(norm_mask * loss_map) - e1) / (torch.sum(norm_mask) + self.eps)

    @staticmethod
    def _column_max(source_norm, target_norm, rows=4096):
        """max over source positions of source_norm @ target_norm without holding the whole
        [b, N2, N2] product (1 GiB per sample at relu1_1): row blocks, running max."""
        best = None
        for r0 in range(0, source_norm.size(1), rows):
            m = torch.bmm(source_norm[:, r0:r0 + rows], target_norm).amax(dim=1)
            best = m if best is None else torch.maximum(best, m)
        return best

    def perceptual_loss(self, flow, layer, norm_mask=None, use_bilinear_sampling=False):
        target_vgg, source_vgg = self.target_vgg[layer], self.source_vgg[layer]
        b, c, h, w = target_vgg.shape
        flow = F.interpolate(flow, [h, w])
        input_sample = self._sample(source_vgg, flow, use_bilinear_sampling)
        if norm_mask is None:
            return self.l1_loss(input_sample, target_vgg)
        norm_mask = F.interpolate(norm_mask, size=(h, w))
        return self.l1_loss(input_sample * norm_mask, target_vgg * norm_mask)

    def bilinear_warp(self, source, flow, view=True):
        return grid_warp(source.contiguous(), flow.contiguous())
