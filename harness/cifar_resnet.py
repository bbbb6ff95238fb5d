"""CIFAR ResNet (depth 6n+2, parameter-free option-A shortcuts): the teacher family of BASELINE config 1, as measurement /
test HARNESS around the cheap-conv blocks (teacher zoos are out of the accelerated path).

A restatement of models/cifar_models/resnet.py:54-126 with the same module names, so `checkpoints/cifar10/resnet44.th`
(tests/golden/cifar_step.npz carries its tensors) loads by key and the config's block names (`layer3.1.conv1`, ...) resolve.
"""
import torch.nn.functional as F
from torch import nn


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
