"""Teacher architectures the step-level parity tests host the kdcc blocks in (see harness/)."""
from torch import nn

from harness.cifar_resnet import CifarResNet  # noqa: F401
from harness.deeplab_wrn38 import ResidualUnit


class ResidualTeacher(nn.Module):
    """stem -> two pre-activation residual units (identity and projected shortcut) -> 1x1 head; the topology of
    tests/golden/residual.npz (built there from the reference's IdentityResidualBlock)."""

    def __init__(self, classes=5):
        super().__init__()
        self.stem = nn.Conv2d(3, 16, 3, padding=1, bias=False)
        self.block1 = ResidualUnit(16, (16, 16), dilation=2)
        self.block2 = ResidualUnit(16, (32, 32), dilation=1)
        self.head = nn.Conv2d(32, classes, 1, bias=False)

    def forward(self, x):
        return self.head(self.block2(self.block1(self.stem(x))))
