import torch
from torch import nn

from oracle import prims


class Conv2d(nn.Conv2d):
    """upstream layers/misc.py Conv2d = nn.Conv2d (+ empty-batch support, unused here)."""


class DeformConv(nn.Module):
    """upstream layers/dcn DeformConv (DCN v1, the original deformable convolution): weight only, no bias.  The arithmetic is
    torchvision.ops.deform_conv2d -- an independent implementation of the same published operator (offset channel 2k = dy,
    2k+1 = dx of kernel tap k, bilinear sampling, zero outside the map)."""

    def __init__(self, in_channels, out_channels, kernel_size, stride=1, padding=0, dilation=1, groups=1, deformable_groups=1, bias=False):
        super().__init__()
        assert not bias and groups == 1 and deformable_groups == 1
        self.weight = nn.Parameter(torch.empty(out_channels, in_channels, kernel_size, kernel_size))
        nn.init.kaiming_uniform_(self.weight, a=1)
        self.stride, self.padding, self.dilation = stride, padding, dilation

    def forward(self, x, offset):
        from torchvision.ops import deform_conv2d
        return deform_conv2d(x, offset, self.weight, None, stride=self.stride, padding=self.padding, dilation=self.dilation)


class DFConv2d(nn.Module):
    """upstream layers/misc.py DFConv2d with with_modulated_dcn=False (the only form dla.py:75-78 builds): a regular conv
    predicts 2*k*k offsets per output pixel (with bias), DeformConv consumes them.  Restated from the upstream semantics."""

    def __init__(self, in_channels, out_channels, with_modulated_dcn=True, kernel_size=3, stride=1, groups=1, dilation=1,
                 deformable_groups=1, bias=False):
        super().__init__()
        if with_modulated_dcn:
            raise NotImplementedError("modulated deformable convolution (not used by siammot/modelling/backbone/dla.py)")
        padding = dilation * (kernel_size - 1) // 2
        self.offset = Conv2d(in_channels, deformable_groups * 2 * kernel_size * kernel_size, kernel_size=kernel_size, stride=stride,
                             padding=padding, groups=1, dilation=dilation)
        nn.init.kaiming_uniform_(self.offset.weight, a=1)
        nn.init.constant_(self.offset.bias, 0.)
        self.conv = DeformConv(in_channels, out_channels, kernel_size, stride=stride, padding=padding, dilation=dilation, groups=groups,
                               deformable_groups=deformable_groups, bias=bias)

    def forward(self, x):
        return self.conv(x, self.offset(x))


class FrozenBatchNorm2d(nn.Module):
    def __init__(self, n):
        super().__init__()
        self.register_buffer("weight", torch.ones(n))
        self.register_buffer("bias", torch.zeros(n))
        self.register_buffer("running_mean", torch.zeros(n))
        self.register_buffer("running_var", torch.ones(n))

    def forward(self, x):
        scale, bias = prims.frozen_bn_scale_bias(self.weight, self.bias, self.running_mean, self.running_var)
        return x * scale.reshape(1, -1, 1, 1) + bias.reshape(1, -1, 1, 1)


class ROIAlign(nn.Module):
    def __init__(self, output_size, spatial_scale, sampling_ratio):
        super().__init__()
        self.output_size = output_size
        self.spatial_scale = spatial_scale
        self.sampling_ratio = sampling_ratio

    def forward(self, input, rois):
        return prims.roi_align_legacy(input, rois, self.spatial_scale, self.output_size[0],
                                      self.output_size[1], self.sampling_ratio)
