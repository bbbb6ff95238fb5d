"""Teacher trunk used as measurement HARNESS around the hot path: DeepLabV3+ on WideResNet38 (137.1 M parameters).

Not part of the accelerated path (SURVEY.md 2.1 marks the teacher zoos OUT: stock PyTorch / cuDNN); it exists so that
`bench.py` can run BASELINE.json's metric as named -- a whole layerwise-KD step at 1024x1024 with the frozen teacher and
the frozen student trunk around the nine kdcc blocks -- and so that the drop-in path (kdcc.DepthwiseStudent +
kdcc.prepare_train_epoch on the reference's own cfg/cityscapes/*.json block names) is exercised on the real topology.

A restatement in this repo's own words of models/deeplabv3/deeplabv3.py:78-162 (`DeepWV3Plus`), :21-75 (ASPP) and
models/encoders/wider_resnet.py:64-182, :267-365 (pre-activation residual units, `WiderResNetA2` with dilation): same
module names, so the reference's `state_dict` keys, config block names (`mod4.block2.convs.conv2`, `aspp.features.1.0`,
...) and parameter counts carry over (tests/test_harness_trunk.py checks keys, shapes and -- in the build container,
where /root/reference exists -- a forward pass against the reference class).
"""
from collections import OrderedDict

import torch
import torch.nn.functional as F
from torch import nn


def bn_act(channels):
    return nn.Sequential(nn.BatchNorm2d(channels), nn.ReLU(inplace=True))


class ResidualUnit(nn.Module):
    """Pre-activation identity-mapping unit.  `widths` of length 2: 3x3 -> 3x3; length 3: 1x1 -> 3x3 -> 1x1 (bottleneck).
    The unit's output tensor is produced by `convs` and then updated IN PLACE with the shortcut: a forward
    hook on the last conv of `convs` therefore sees output + shortcut."""

    def __init__(self, cin, widths, stride=1, dilation=1, dropout=None):
        super().__init__()
        self.bn1 = bn_act(cin)
        if len(widths) == 2:
            layers = [("conv1", nn.Conv2d(cin, widths[0], 3, stride=stride, padding=dilation, dilation=dilation, bias=False)),
                      ("bn2", bn_act(widths[0])),
                      ("conv2", nn.Conv2d(widths[0], widths[1], 3, padding=dilation, dilation=dilation, bias=False))]
            if dropout:
                layers.insert(2, ("dropout", nn.Dropout2d(dropout)))
        else:
            layers = [("conv1", nn.Conv2d(cin, widths[0], 1, stride=stride, bias=False)),
                      ("bn2", bn_act(widths[0])),
                      ("conv2", nn.Conv2d(widths[0], widths[1], 3, padding=dilation, dilation=dilation, bias=False)),
                      ("bn3", bn_act(widths[1])),
                      ("conv3", nn.Conv2d(widths[1], widths[2], 1, bias=False))]
            if dropout:
                layers.insert(4, ("dropout", nn.Dropout2d(dropout)))
        self.convs = nn.Sequential(OrderedDict(layers))
        if stride != 1 or cin != widths[-1]:
            self.proj_conv = nn.Conv2d(cin, widths[-1], 1, stride=stride, bias=False)

    def forward(self, x):
        if hasattr(self, "proj_conv"):
            pre = self.bn1(x)
            shortcut = self.proj_conv(pre)
        else:
            shortcut = x.clone()   # bn1's ReLU works in place
            pre = self.bn1(x)
        out = self.convs(pre)
        out.add_(shortcut)
        return out


WRN38_WIDTHS = [(128, 128), (256, 256), (512, 512), (512, 1024), (512, 1024, 2048), (1024, 2048, 4096)]
WRN38_DEPTHS = [3, 3, 6, 3, 1, 1]


def _stage(cin, widths, depth, stride, dilation, dropout):
    units = []
    for i in range(depth):
        units.append(("block%d" % (i + 1), ResidualUnit(cin, widths, stride if i == 0 else 1, dilation, dropout)))
        cin = widths[-1]
    return nn.Sequential(OrderedDict(units)), cin


class ASPP(nn.Module):
    """1x1 + three dilated 3x3 branches + image pooling, concatenated (image feature first)."""

    def __init__(self, cin, width=256, rates=(12, 24, 36)):
        super().__init__()
        branch = lambda k, r: nn.Sequential(nn.Conv2d(cin, width, k, padding=r if k == 3 else 0, dilation=r if k == 3 else 1, bias=False),
                                            nn.BatchNorm2d(width), nn.ReLU(inplace=True))
        self.features = nn.ModuleList([branch(1, 1)] + [branch(3, r) for r in rates])
        self.img_pooling = nn.AdaptiveAvgPool2d(1)
        self.img_conv = branch(1, 1)

    def forward(self, x):
        img = self.img_conv(self.img_pooling(x))
        img = F.interpolate(img, size=x.shape[2:], mode="bilinear", align_corners=True)
        return torch.cat([img] + [f(x) for f in self.features], 1)


class DeepWV3Plus(nn.Module):
    """DeepLabV3+ head on the dilated WideResNet38 trunk: output stride 8, decoder skip from mod2 (stride 2)."""

    def __init__(self, num_classes=19):
        super().__init__()
        self.mod1 = nn.Sequential(OrderedDict([("conv1", nn.Conv2d(3, 64, 3, padding=1, bias=False))]))
        cin = 64
        #          stride dilation dropout
        plan = [(1, 1, None), (1, 1, None), (2, 1, None), (1, 2, None), (1, 4, 0.3), (1, 4, 0.5)]
        for i, (widths, depth, (stride, dil, drop)) in enumerate(zip(WRN38_WIDTHS, WRN38_DEPTHS, plan)):
            stage, cin = _stage(cin, widths, depth, stride, dil, drop)
            if i < 2:
                setattr(self, "pool%d" % (i + 2), nn.MaxPool2d(3, stride=2, padding=1))
            setattr(self, "mod%d" % (i + 2), stage)
        self.aspp = ASPP(4096, 256)
        self.bot_fine = nn.Conv2d(128, 48, 1, bias=False)
        self.bot_aspp = nn.Conv2d(1280, 256, 1, bias=False)
        self.final = nn.Sequential(nn.Conv2d(256 + 48, 256, 3, padding=1, bias=False), nn.BatchNorm2d(256), nn.ReLU(inplace=True),
                                   nn.Conv2d(256, 256, 3, padding=1, bias=False), nn.BatchNorm2d(256), nn.ReLU(inplace=True),
                                   nn.Conv2d(256, num_classes, 1, bias=False))

    def forward(self, x):
        size = x.shape[2:]
        x = self.mod1(x)
        m2 = self.mod2(self.pool2(x))
        x = self.mod3(self.pool3(m2))
        x = self.mod7(self.mod6(self.mod5(self.mod4(x))))
        x = self.bot_aspp(self.aspp(x))
        x = F.interpolate(x, size=m2.shape[2:], mode="bilinear", align_corners=True)
        x = self.final(torch.cat([self.bot_fine(m2), x], 1))
        return F.interpolate(x, size=size, mode="bilinear", align_corners=True)


# cfg/cityscapes/51M_deeplab_all.json:117-160 -- the nine convs the shipped 51M plan replaces, in network order
PRUNING_51M = {
    "args": {"dilation": 5, "padding": 20, "kernel_size": 9},
    "names": ["mod4.block2.convs.conv2", "mod4.block3.convs.conv1", "mod4.block3.convs.conv2", "mod4.block4.convs.conv2",
              "mod4.block6.convs.conv2", "mod7.block1.convs.conv2", "aspp.features.1.0", "aspp.features.2.0", "aspp.features.3.0"],
}


def pruning_section(plan=PRUNING_51M, epoch=1):
    """The config's "pruning" section for `kdcc.prepare_train_epoch`: every listed conv is replaced, hinted and unfrozen at
    `epoch`, exactly as the shipped cfg does."""
    entries = [{"name": n, "epoch": epoch} for n in plan["names"]]
    return {"args": dict(plan["args"]), "pruning_plan": [dict(e) for e in entries], "hint": [dict(e) for e in entries],
            "unfreeze": [dict(e) for e in entries]}
