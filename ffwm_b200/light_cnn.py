"""Host-side mirror of the reference's `lightcnn/light_cnn.py`: LightCNN with
max-feature-map (MFM) activations, used frozen as the identity-preserving feature extractor
(`models/losses.py:76-112`).  Same class / factory names and `state_dict()` keys
(`conv1.filter.weight`, `block2.1.conv1.filter.bias`, `fc.filter.weight`, ...), so
`LightCNN_29Layers_checkpoint.pth` loads unchanged.

    mfm                light_cnn.py:13-26    conv/linear to 2*C channels, split, elementwise max
    group / resblock   light_cnn.py:29-54
    network_29layers   light_cnn.py:82-129   returns (logits, fc256, pooled 128x8x8)
"""
import os

import torch
import torch.nn as nn
import torch.nn.functional as F

from . import ops
from .conv import Conv2d
from .pool import MaxPool2d, max_pool2x2

# max-feature-map as one kernel per direction (csrc/mfm.cu): bit-exact vs torch.max on a B200 (tests/test_mfm_gpu.py), train step
# 73.4 -> 72.0 ms (profiles/r02b_switches.txt).  FFWM_FUSED_MFM=0 restores the torch ops for A/B runs.
FUSED_MFM = os.environ.get("FFWM_FUSED_MFM", "1") == "1"


class MFMFunction(torch.autograd.Function):
    """max over the two channel halves as one kernel per direction (PyTorch: 1 forward + 8 backward kernels)."""

    @staticmethod
    def forward(ctx, y):
        y = y.contiguous()
        out = y.new_empty((y.size(0), y.size(1) // 2) + tuple(y.shape[2:]))
        ops.mfm_forward(y, out)
        ctx.save_for_backward(y)
        return out

    @staticmethod
    def backward(ctx, grad_out):
        (y,) = ctx.saved_tensors
        gy = torch.empty_like(y)
        ops.mfm_backward(y, grad_out.contiguous(), gy)
        return gy


class mfm(nn.Module):
    def __init__(self, in_channels, out_channels, kernel_size=3, stride=1, padding=1, type=1):
        super().__init__()
        self.out_channels = out_channels
        if type == 1:
            self.filter = Conv2d(in_channels, 2 * out_channels, kernel_size=kernel_size, stride=stride, padding=padding)
        else:
            self.filter = nn.Linear(in_channels, 2 * out_channels)

    def forward(self, x):
        y = self.filter(x)
        if FUSED_MFM and y.is_cuda and y.dtype == torch.float32:
            return MFMFunction.apply(y)
        a, b = y.split(self.out_channels, 1)
        return torch.max(a, b)


class group(nn.Module):
    """1x1 MFM (channel mixing) followed by a kxk MFM."""

    def __init__(self, in_channels, out_channels, kernel_size, stride, padding, mid_channels=None):
        super().__init__()
        mid = in_channels if mid_channels is None else mid_channels
        self.conv_a = mfm(in_channels, mid, 1, 1, 0)
        self.conv = mfm(mid, out_channels, kernel_size, stride, padding)

    def forward(self, x):
        return self.conv(self.conv_a(x))


class resblock(nn.Module):
    def __init__(self, in_channels, out_channels):
        super().__init__()
        self.conv1 = mfm(in_channels, out_channels, kernel_size=3, stride=1, padding=1)
        self.conv2 = mfm(in_channels, out_channels, kernel_size=3, stride=1, padding=1)

    def forward(self, x):
        return self.conv2(self.conv1(x)) + x


def _stack(block, count, cin, cout):
    return nn.Sequential(*[block(cin, cout) for _ in range(count)])


def _pool():
    return MaxPool2d(kernel_size=2, stride=2, ceil_mode=True)


class network_9layers(nn.Module):
    """light_cnn.py:57-80 (not used by FFWM; kept because the factory is public)."""

    def __init__(self, num_classes=79077):
        super().__init__()
        self.features = nn.Sequential(
            mfm(1, 48, 5, 1, 2), _pool(), group(48, 96, 3, 1, 1), _pool(), group(96, 192, 3, 1, 1), _pool(),
            group(192, 128, 3, 1, 1), group(128, 128, 3, 1, 1), _pool())
        self.fc1 = mfm(8 * 8 * 128, 256, type=0)
        self.fc2 = nn.Linear(256, num_classes)

    def forward(self, x):
        x = self.fc1(self.features(x).flatten(1))
        x = F.dropout(x, training=self.training)
        return self.fc2(x), x


class network_29layers(nn.Module):
    # (stage width in, stage width out) of the four residual stages and the MFM group after each
    STAGES = ((48, 96), (96, 192), (192, 128), (128, 128))

    def __init__(self, block, layers, num_classes=79077):
        super().__init__()
        self.conv1 = mfm(1, 48, 5, 1, 2)
        self.pool1 = _pool()
        self.block1 = _stack(block, layers[0], 48, 48)
        self.group1 = group(48, 96, 3, 1, 1)
        self.pool2 = _pool()
        self.block2 = _stack(block, layers[1], 96, 96)
        self.group2 = group(96, 192, 3, 1, 1)
        self.pool3 = _pool()
        self.block3 = _stack(block, layers[2], 192, 192)
        self.group3 = group(192, 128, 3, 1, 1)
        self.block4 = _stack(block, layers[3], 128, 128)
        self.group4 = group(128, 128, 3, 1, 1)
        self.pool4 = _pool()
        self.fc = mfm(8 * 8 * 128, 256, type=0)
        self.fc2 = nn.Linear(256, num_classes)

    def _make_layer(self, block, num_blocks, in_channels, out_channels):
        return _stack(block, num_blocks, in_channels, out_channels)

    def forward(self, x):
        x = self.pool1(self.conv1(x))
        x = self.pool2(self.group1(self.block1(x)))
        x = self.pool3(self.group2(self.block2(x)))
        x = self.group3(self.block3(x))
        p = self.pool4(self.group4(self.block4(x)))
        fc = F.dropout(self.fc(p.flatten(1)), training=self.training)    # the POST-dropout features are returned (:124-127)
        return self.fc2(fc), fc, p


class network_29layers_v2(nn.Module):
    """light_cnn.py:131-175: max+avg pooling, plain linear fc."""

    def __init__(self, block, layers, num_classes=80013):
        super().__init__()
        self.conv1 = mfm(1, 48, 5, 1, 2)
        self.block1 = _stack(block, layers[0], 48, 48)
        self.group1 = group(48, 96, 3, 1, 1)
        self.block2 = _stack(block, layers[1], 96, 96)
        self.group2 = group(96, 192, 3, 1, 1)
        self.block3 = _stack(block, layers[2], 192, 192)
        self.group3 = group(192, 128, 3, 1, 1)
        self.block4 = _stack(block, layers[3], 128, 128)
        self.group4 = group(128, 128, 3, 1, 1)
        self.fc = nn.Linear(8 * 8 * 128, 256)
        self.fc2 = nn.Linear(256, num_classes, bias=False)

    @staticmethod
    def _mixpool(x):
        return max_pool2x2(x) + F.avg_pool2d(x, 2)

    def forward(self, x):
        x = self._mixpool(self.conv1(x))
        x = self._mixpool(self.group1(self.block1(x)))
        x = self._mixpool(self.group2(self.block2(x)))
        x = self.group3(self.block3(x))
        p = self._mixpool(self.group4(self.block4(x)))
        fc = self.fc(p.flatten(1))
        out = self.fc2(F.dropout(fc, training=self.training))
        return out, fc, p


def LightCNN_9Layers(**kwargs):
    return network_9layers(**kwargs)


def LightCNN_29Layers(**kwargs):
    return network_29layers(resblock, [1, 2, 3, 4], **kwargs)


def LightCNN_29Layers_v2(**kwargs):
    return network_29layers_v2(resblock, [1, 2, 3, 4], **kwargs)
