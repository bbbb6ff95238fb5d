"""Teacher architectures the step-level parity tests host the kdcc blocks in (test infrastructure, not product).

The trunk around the replaced convolutions is out of the hot path's scope (stock PyTorch); the golden
fixtures tests/golden/{residual,cifar_step}.npz were produced by the reference's own model classes, which do not exist on
the GPU box.  These are minimal restatements with the same module names -- so the frozen `state_dict` loads by key --
and the same arithmetic:

  PreActResidualUnit   models/encoders/wider_resnet.py:64-182 (two-conv form): bn1 -> [proj_conv] -> conv1, bn2, conv2,
                       then the IN-PLACE `out.add_(shortcut)` on the tensor the forward hook of conv2 stored.
  CifarResNet          models/cifar_models/resnet.py:54-126: CIFAR ResNet with the parameter-free option-A shortcut
                       (stride-2 subsample + zero channel padding), depth 6n+2.
"""
from collections import OrderedDict

import torch
import torch.nn.functional as F
from torch import nn


def _bn_act(c):
    return nn.Sequential(nn.BatchNorm2d(c), nn.ReLU(inplace=True))


class PreActResidualUnit(nn.Module):
    def __init__(self, cin, mid, cout, dilation=1):
        super().__init__()
        self.bn1 = _bn_act(cin)
        self.convs = nn.Sequential(OrderedDict([
            ("conv1", nn.Conv2d(cin, mid, 3, padding=dilation, dilation=dilation, bias=False)),
            ("bn2", _bn_act(mid)),
            ("conv2", nn.Conv2d(mid, cout, 3, padding=dilation, dilation=dilation, bias=False))]))
        if cin != cout:
            self.proj_conv = nn.Conv2d(cin, cout, 1, bias=False)

    def forward(self, x):
        if hasattr(self, "proj_conv"):
            pre = self.bn1(x)
            shortcut = self.proj_conv(pre)
        else:
            shortcut = x.clone()      # bn1's ReLU is in place: the shortcut must be taken before it
            pre = self.bn1(x)
        out = self.convs(pre)
        out.add_(shortcut)            # mutates the tensor a forward hook on convs.conv2 has stored
        return out


class ResidualTeacher(nn.Module):
    """stem -> two residual units (identity and projected shortcut) -> 1x1 head; tests/golden/residual.npz."""

    def __init__(self, classes=5):
        super().__init__()
        self.stem = nn.Conv2d(3, 16, 3, padding=1, bias=False)
        self.block1 = PreActResidualUnit(16, 16, 16, dilation=2)
        self.block2 = PreActResidualUnit(16, 32, 32, dilation=1)
        self.head = nn.Conv2d(32, classes, 1, bias=False)

    def forward(self, x):
        return self.head(self.block2(self.block1(self.stem(x))))


class _PadShortcut(nn.Module):
    """Option A of the CIFAR ResNet paper: every second pixel, zero channels on both sides."""

    def __init__(self, pad):
        super().__init__()
        self.pad = pad

    def forward(self, x):
        return F.pad(x[:, :, ::2, ::2], (0, 0, 0, 0, self.pad, self.pad))


class _CifarBlock(nn.Module):
    def __init__(self, cin, cout, stride):
        super().__init__()
        self.conv1 = nn.Conv2d(cin, cout, 3, stride=stride, padding=1, bias=False)
        self.bn1 = nn.BatchNorm2d(cout)
        self.conv2 = nn.Conv2d(cout, cout, 3, padding=1, bias=False)
        self.bn2 = nn.BatchNorm2d(cout)
        self.shortcut = _PadShortcut(cout // 4) if (stride != 1 or cin != cout) else nn.Sequential()

    def forward(self, x):
        y = F.relu(self.bn1(self.conv1(x)))
        y = self.bn2(self.conv2(y))
        y += self.shortcut(x)
        return F.relu(y)


class CifarResNet(nn.Module):
    def __init__(self, n=7, classes=10):   # n = 7 -> ResNet44
        super().__init__()
        self.conv1 = nn.Conv2d(3, 16, 3, padding=1, bias=False)
        self.bn1 = nn.BatchNorm2d(16)
        cin, stages = 16, []
        for cout, stride in ((16, 1), (32, 2), (64, 2)):
            blocks = []
            for i in range(n):
                blocks.append(_CifarBlock(cin, cout, stride if i == 0 else 1))
                cin = cout
            stages.append(nn.Sequential(*blocks))
        self.layer1, self.layer2, self.layer3 = stages
        self.linear = nn.Linear(64, classes)

    def forward(self, x):
        y = F.relu(self.bn1(self.conv1(x)))
        y = self.layer3(self.layer2(self.layer1(y)))
        y = F.avg_pool2d(y, y.size(3)).flatten(1)
        return self.linear(y)
