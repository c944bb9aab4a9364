import torch
from torch import nn

from oracle import prims
from maskrcnn_benchmark.structures.bounding_box import BoxList


class BufferList(nn.Module):
    def __init__(self, buffers):
        super().__init__()
        for i, b in enumerate(buffers):
            self.register_buffer(str(i), b)

    def __len__(self):
        return len(self._buffers)

    def __iter__(self):
        return iter(self._buffers.values())


class AnchorGenerator(nn.Module):
    def __init__(self, sizes, aspect_ratios, anchor_strides, straddle_thresh=0):
        super().__init__()
        assert len(anchor_strides) == len(sizes), "FPN mode: one size per stride"
        cells = [prims.cell_anchors(st, sz if isinstance(sz, (tuple, list)) else (sz,), aspect_ratios)
                 for st, sz in zip(anchor_strides, sizes)]
        self.strides = anchor_strides
        self.cell_anchors = BufferList(cells)
        self.straddle_thresh = straddle_thresh

    def num_anchors_per_location(self):
        return [len(c) for c in self.cell_anchors]

    def forward(self, image_list, feature_maps):
        per_level = [prims.grid_anchors(c, st, fm.shape[-2], fm.shape[-1])
                     for fm, st, c in zip(feature_maps, self.strides, self.cell_anchors)]
        anchors = []
        for (h, w) in image_list.image_sizes:
            anchors.append([BoxList(a, (w, h), mode="xyxy") for a in per_level])
        return anchors


def make_anchor_generator(config):
    return AnchorGenerator(config.MODEL.RPN.ANCHOR_SIZES, config.MODEL.RPN.ASPECT_RATIOS,
                           config.MODEL.RPN.ANCHOR_STRIDE, config.MODEL.RPN.STRADDLE_THRESH)
