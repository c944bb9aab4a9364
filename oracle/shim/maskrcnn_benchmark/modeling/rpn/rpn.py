import torch
import torch.nn.functional as F
from torch import nn

from maskrcnn_benchmark.modeling import registry
from maskrcnn_benchmark.modeling.box_coder import BoxCoder
from .anchor_generator import make_anchor_generator
from .inference import make_rpn_postprocessor  # bound at import time: rpn_patch must run first


@registry.RPN_HEADS.register("SingleConvRPNHead")
class RPNHead(nn.Module):
    def __init__(self, cfg, in_channels, num_anchors):
        super().__init__()
        self.conv = nn.Conv2d(in_channels, in_channels, kernel_size=3, stride=1, padding=1)
        self.cls_logits = nn.Conv2d(in_channels, num_anchors, kernel_size=1, stride=1)
        self.bbox_pred = nn.Conv2d(in_channels, num_anchors * 4, kernel_size=1, stride=1)
        for l in [self.conv, self.cls_logits, self.bbox_pred]:
            nn.init.normal_(l.weight, std=0.01)
            nn.init.constant_(l.bias, 0)

    def forward(self, x):
        logits, bbox_reg = [], []
        for feature in x:
            t = F.relu(self.conv(feature))
            logits.append(self.cls_logits(t))
            bbox_reg.append(self.bbox_pred(t))
        return logits, bbox_reg


class RPNModule(nn.Module):
    def __init__(self, cfg, in_channels):
        super().__init__()
        self.cfg = cfg.clone()
        anchor_generator = make_anchor_generator(cfg)
        head = registry.RPN_HEADS[cfg.MODEL.RPN.RPN_HEAD](
            cfg, in_channels, anchor_generator.num_anchors_per_location()[0])
        rpn_box_coder = BoxCoder(weights=(1.0, 1.0, 1.0, 1.0))
        self.anchor_generator = anchor_generator
        self.head = head
        self.box_selector_test = make_rpn_postprocessor(cfg, rpn_box_coder, is_train=False)

    def forward(self, images, features, targets=None):
        assert not self.training, "stand-in supports inference only"
        objectness, rpn_box_regression = self.head(features)
        anchors = self.anchor_generator(images, features)
        boxes = self.box_selector_test(anchors, objectness, rpn_box_regression)
        if self.cfg.MODEL.RPN_ONLY:
            inds = [b.get_field("objectness").sort(descending=True)[1] for b in boxes]
            boxes = [b[i] for b, i in zip(boxes, inds)]
        return boxes, {}


def build_rpn(cfg, in_channels):
    return RPNModule(cfg, in_channels)
