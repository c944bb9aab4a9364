"""Upstream modeling/backbone/backbone.py, the FPN ResNet builders (restated).  ``fpn_module.FPN`` is looked up at call
time, i.e. it is the class siammot/operator_patch/fpn_patch.py:65 installed."""
from collections import OrderedDict

from torch import nn

from maskrcnn_benchmark.modeling import registry
from maskrcnn_benchmark.modeling.make_layers import conv_with_kaiming_uniform

from . import fpn as fpn_module
from . import resnet


@registry.BACKBONES.register("R-50-FPN")
@registry.BACKBONES.register("R-101-FPN")
def build_resnet_fpn_backbone(cfg):
    body = resnet.ResNet(cfg)
    c2 = cfg.MODEL.RESNETS.RES2_OUT_CHANNELS
    out_channels = cfg.MODEL.RESNETS.BACKBONE_OUT_CHANNELS
    fpn = fpn_module.FPN(in_channels_list=[c2, c2 * 2, c2 * 4, c2 * 8], out_channels=out_channels,
                         conv_block=conv_with_kaiming_uniform(cfg.MODEL.FPN.USE_GN, cfg.MODEL.FPN.USE_RELU),
                         top_blocks=fpn_module.LastLevelMaxPool())
    model = nn.Sequential(OrderedDict([("body", body), ("fpn", fpn)]))
    model.out_channels = out_channels
    return model
