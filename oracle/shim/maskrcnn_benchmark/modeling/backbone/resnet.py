"""Upstream maskrcnn_benchmark/modeling/backbone/resnet.py, inference subset, restated: ResNet-50 stages 2..5 returned
for the FPN ("R-50-FPN"), FrozenBatchNorm stem + bottleneck blocks, stride on the first 1x1 (STRIDE_IN_1X1, the Detectron
convention).  TEST INFRASTRUCTURE (see the package docstring).  The un-vendored upstream is restated from its published
semantics; tests/test_oracle_resnet.py cross-checks this file against torchvision's ResNet-50 (strides moved to the 1x1)."""
from collections import namedtuple

import torch.nn.functional as F
from torch import nn

from maskrcnn_benchmark.layers import Conv2d, FrozenBatchNorm2d

StageSpec = namedtuple("StageSpec", ["index", "block_count", "return_features"])

# (stage index, residual blocks, returned to the FPN)
_R50_FPN = tuple(StageSpec(index=i, block_count=c, return_features=True) for i, c in ((1, 3), (2, 4), (3, 6), (4, 3)))
_R101_FPN = tuple(StageSpec(index=i, block_count=c, return_features=True) for i, c in ((1, 3), (2, 4), (3, 23), (4, 3)))
_STAGE_SPECS = {"R-50-FPN": _R50_FPN, "R-101-FPN": _R101_FPN}


class Bottleneck(nn.Module):
    def __init__(self, in_channels, bottleneck_channels, out_channels, num_groups, stride_in_1x1, stride):
        super().__init__()
        self.downsample = None
        if in_channels != out_channels:
            self.downsample = nn.Sequential(Conv2d(in_channels, out_channels, kernel_size=1, stride=stride, bias=False),
                                            FrozenBatchNorm2d(out_channels))
        s1, s3 = (stride, 1) if stride_in_1x1 else (1, stride)
        self.conv1 = Conv2d(in_channels, bottleneck_channels, kernel_size=1, stride=s1, bias=False)
        self.bn1 = FrozenBatchNorm2d(bottleneck_channels)
        self.conv2 = Conv2d(bottleneck_channels, bottleneck_channels, kernel_size=3, stride=s3, padding=1, bias=False,
                            groups=num_groups)
        self.bn2 = FrozenBatchNorm2d(bottleneck_channels)
        self.conv3 = Conv2d(bottleneck_channels, out_channels, kernel_size=1, bias=False)
        self.bn3 = FrozenBatchNorm2d(out_channels)

    def forward(self, x):
        identity = x
        out = F.relu_(self.bn1(self.conv1(x)))
        out = F.relu_(self.bn2(self.conv2(out)))
        out = self.bn3(self.conv3(out))
        if self.downsample is not None:
            identity = self.downsample(x)
        out += identity
        return F.relu_(out)


class Stem(nn.Module):
    def __init__(self, cfg):
        super().__init__()
        out_channels = cfg.MODEL.RESNETS.STEM_OUT_CHANNELS
        self.conv1 = Conv2d(3, out_channels, kernel_size=7, stride=2, padding=3, bias=False)
        self.bn1 = FrozenBatchNorm2d(out_channels)

    def forward(self, x):
        x = F.relu_(self.bn1(self.conv1(x)))
        return F.max_pool2d(x, kernel_size=3, stride=2, padding=1)


class ResNet(nn.Module):
    def __init__(self, cfg):
        super().__init__()
        R = cfg.MODEL.RESNETS
        assert R.STEM_FUNC == "StemWithFixedBatchNorm" and R.TRANS_FUNC == "BottleneckWithFixedBatchNorm"
        assert R.RES5_DILATION == 1 and not any(R.STAGE_WITH_DCN)
        self.stem = Stem(cfg)
        in_channels = R.STEM_OUT_CHANNELS
        self.stages, self.return_features = [], {}
        for spec in _STAGE_SPECS[cfg.MODEL.BACKBONE.CONV_BODY]:
            name = "layer%d" % spec.index
            factor = 2 ** (spec.index - 1)
            bottleneck = R.NUM_GROUPS * R.WIDTH_PER_GROUP * factor
            out_channels = R.RES2_OUT_CHANNELS * factor
            blocks, stride = [], int(spec.index > 1) + 1
            for _ in range(spec.block_count):
                blocks.append(Bottleneck(in_channels, bottleneck, out_channels, R.NUM_GROUPS, R.STRIDE_IN_1X1, stride))
                stride, in_channels = 1, out_channels
            self.add_module(name, nn.Sequential(*blocks))
            self.stages.append(name)
            self.return_features[name] = spec.return_features

    def forward(self, x):
        outputs = []
        x = self.stem(x)
        for name in self.stages:
            x = getattr(self, name)(x)
            if self.return_features[name]:
                outputs.append(x)
        return outputs
